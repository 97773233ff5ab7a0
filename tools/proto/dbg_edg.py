import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import simlib, sigops_oracle as o, unit_checks as uc
import wgpu_sigops_b200 as w
lib = w.load()
U = simlib.UnitRunner(lib.sigops_test_unit, lib.sigops_test_unit_shape)
rng = random.Random(1)
px, py = uc._affine_ed(o.ed_mul(12345, o.ED_B))
for k in (0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 17, 255, 2**64 + 5):
    a = U.run("ED_MULPT", [(k, px, py)])[0]
    b = U.run("ED_GROUP_MULPT", [(k, px, py)])[0]
    e = uc._affine_ed(o.ed_mul(k, (px, py, 1, px * py % o.ED_P)))
    print(k, "thread ok" if (simlib.from_words(a[:8]), simlib.from_words(a[8:16])) == e else "thread BAD",
          "group ok" if (simlib.from_words(b[:8]), simlib.from_words(b[8:16])) == e else "group BAD")
items = [(rng.getrandbits(250), px, py) for _ in range(40)]
a = U.run("ED_MULPT", items); b = U.run("ED_GROUP_MULPT", items)
print("40 items equal rows:", (a == b).all(axis=1).astype(int))
