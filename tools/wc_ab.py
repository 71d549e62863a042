"""A/B two builds of the fused warp+conv kernel in one process launch each (same box)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[1:]:
    env = dict(os.environ, FUSED_ONLY="1")
    if lib != "default":
        env["DVC_B200_LIB"] = os.path.join(ROOT, "deepvideocodec_b200", lib.split(":")[0])
        if lib.endswith(":old"):
            env["OLD_ABI"] = "1"
    for rep in range(1):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "warp_conv_bench.py")], env=env,
                             capture_output=True, text=True)
        try:
            r = json.loads(out.stdout[out.stdout.index("{"):])
            print(lib, rep, "fused_us=%.1f" % r["fused_us"], "bit_exact", r["warp_bit_exact"], flush=True)
        except Exception as e:
            print(lib, rep, "FAILED", out.stderr[-400:], flush=True)
