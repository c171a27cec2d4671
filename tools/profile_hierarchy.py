"""A hierarchy frame (BASELINE configs[2]: 2^22 entities, depth 8) for ncu captures of prepare_level / emit_kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tuber_legacy_b200 import scenes
import util

sc = scenes.config3()
ctx = util.make_ctx(sc, cuda_graphs=0)
for _ in range(4):
    info = ctx.frame()
print(info.num_instances, info.sort_passes, info.hierarchy_levels, info.kernel_launches)
ctx.close()
