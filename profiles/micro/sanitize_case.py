import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from rednoise_b200 import capi
from tests import oracle_lib
o = oracle_lib.load_oracle()
with capi.Context(0) as ctx:
    for depth, seed in ((6, 3), (10, 12345)):
        p = capi.default_params(seed=seed, depth=depth, apply_volcano=1)
        got = ctx.terrain2d_host(p)
        po = oracle_lib.params(seed=seed, depth=depth, apply_volcano=1)
        want = o.terrain2d(po)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), depth
        cells = ctx.cells_host(p)
        assert cells.tobytes() == o.cells(po, want).tobytes()
    print("sanitize case ok")
