import sys, os, subprocess, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys; sys.path.insert(0, %r)
import bsw_b200 as B
ctx = B.Context(); t = B.synth_tasks("cfg2_150bp", 1000000); p = B.make_params()
r = ctx.resident(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
best = min(r.run()[0] for _ in range(6)); ms, cells, nl = r.run()
print("%%.3f ms %%.1f GCUPS launches=%%d" %% (best, cells / best * 1e-6, nl))
''' % root
for side in (2, 3, 5):
    for pct in (115, 130, 160, 200):
        env = dict(os.environ, BSW_SIDE_STREAMS=str(side), BSW_BUCKET_PCT=str(pct))
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip()
        print("side", side, "bucket_pct", pct, "->", out, flush=True)
