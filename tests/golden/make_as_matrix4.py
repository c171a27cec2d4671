"""Writes tests/golden/as_matrix4.json from the vectors SURVEY.md §8c records.

Provenance: the reference (Rust) cannot be built in this image, and `Transform::as_matrix4`
(crates/tuber-core/src/transform.rs:28-38) has no reference test. The survey derived these
vectors independently of oracle/ — a numpy float32 emulation of the four-product chain with
this image's glibc sinf/cosf — and recorded them as hex in SURVEY.md §8c (GV1-GV3). This
script only transcribes them; tests/test_oracle_math.py checks the C++ oracle reproduces them
bit for bit."""
import json
import os

CASES = [
    ("GV1", [10, 20, 0, 0, 0, 0.5, 0, 0, 0, 2, 3, 1],
     "3fe0a940 bf757744 0 41a00000 3fb81973 40287ef0 0 42700000 0 0 3f800000 0 0 0 0 3f800000"),
    ("GV2", [1.5, -2.25, 0.75, 0.1, 0.2, 0.3, 4, 5, 6, 1, 2, 3],
     "3f6fb0eb be8cd95d 3e5f9752 3fe8fc60 3f144a51 3ff4d846 bd97603a c0be027a bf1893fc 3e964996 403b3b92 40674090 0 0 0 3f800000"),
    ("GV3", [100, -50, 3, 0, 0, 2.5, 16, 16, 0, 1.5, 1.5, 1],
     "bf99d1d0 bf65d036 0 434f973e 3f65d036 bf99d1d0 0 c2388b26 0 0 3f800000 40400000 0 0 0 3f800000"),
    ("default", [0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1],
     "3f800000 0 0 0 0 3f800000 0 0 0 0 3f800000 0 0 0 0 3f800000"),
]

out = {"source": "SURVEY.md section 8c, float32 emulation of transform.rs:28-38 (row-major values[0..16])",
       "cases": [{"name": n, "transform": t, "matrix_hex": h.split()} for n, t, h in CASES]}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "as_matrix4.json"), "w"), indent=1)
