"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. Nothing under tuber_legacy_b200/ does.
"""
