"""A/B builds of the fused warp+conv kernel across flow regimes (same box)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:]
for flow in ("smooth", "rigid", "iid"):
    for lib in libs:
        env = dict(os.environ, FUSED_ONLY="1", FLOW=flow)
        if lib != "default":
            env["DVC_B200_LIB"] = os.path.join(ROOT, "deepvideocodec_b200", lib)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "warp_conv_bench.py")], env=env,
                             capture_output=True, text=True)
        try:
            r = json.loads(out.stdout[out.stdout.index("{"):])
            print(flow, lib, "fused_us=%.1f" % r["fused_us"], "bit_exact", r["warp_bit_exact"], "conv_err", r["conv_vs_cudnn_tf32_max_abs"], flush=True)
        except Exception as e:
            print(flow, lib, "FAILED", out.stderr[-300:], flush=True)
