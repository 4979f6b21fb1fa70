"""Development: time one bench step with several experimental builds of the library (QTOS_LIB override).
usage: python tools/variant_bench.py [n=4096] lib1.so lib2.so ..."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if sys.argv[1].isdigit() else "4096"
libs = [a for a in sys.argv[1:] if not a.isdigit()]
for lib in libs:
    env = dict(os.environ)
    if lib != "default":
        env["QTOS_LIB"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "profile_step.py"), n, "3"], env=env, capture_output=True, text=True)
    print("==", lib, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
