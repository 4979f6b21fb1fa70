"""Development: solve the same 512 bench windows with two builds of the library (QTOS_LIB override) and compare
statuses, iteration counts and node values.  usage: python tools/variant_equal.py a.so b.so"""
import os, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 3 and sys.argv[1] != "--child":
    import numpy as np
    outs = []
    for lib in sys.argv[1:3]:
        f = tempfile.mktemp(suffix=".npz")
        env = dict(os.environ, QTOS_LIB=os.path.abspath(lib))
        subprocess.check_call([sys.executable, __file__, "--child", f], env=env)
        outs.append(np.load(f))
    a, b = outs
    print("status equal:", bool((a["status"] == b["status"]).all()), " iters equal:", bool((a["iters"] == b["iters"]).all()),
          " max |dx|: %.3e" % np.abs(a["x"] - b["x"]).max(), " bit-identical:", bool(np.array_equal(a["x"], b["x"])))
else:
    sys.path.insert(0, root)
    import numpy as np
    import qtos_b200 as Q
    from bench import build_workload, COMBO, DURATION
    grid, res, p = build_workload(512)
    S = Q.Solver(Q.default_shape(COMBO, DURATION), max_batch=512)
    p["hf_id"] = S.upload_heightfield(grid, res)
    r, x, _ = S.solve(p)
    np.savez(sys.argv[2], status=r["status"], iters=r["iters"], x=x)
