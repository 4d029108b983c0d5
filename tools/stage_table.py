"""print value + the GEMM rows of the stage table of a bench JSON line (stdin)"""
import json, sys
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
keys = ["l0.pool_gemm", "l0.out_gemm", "l1.dx_gemm", "l0.dneigh_gemm", "l1.pool_gemm", "l0.dW_pool", "l0.dW_group", "l0.pool_bwd", "l0.segmax", "gather"]
print("%s %s: %.4f ms/step (%.3f M/s), e2e %.3f M/s | " % (sys.argv[1] if len(sys.argv) > 1 else "", d["dtype"], d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6) +
      " ".join("%s=%.1fus" % (k.replace("_gemm", ""), 1e3 * d["stages"][k]["ms"]) for k in keys if k in d.get("stages", {})))
