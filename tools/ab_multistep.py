import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
for rep in range(2):
    for mode in ("multi", "single"):
        if mode == "single": os.environ["OGL_NO_MULTISTEP"] = "1"
        else: os.environ.pop("OGL_NO_MULTISTEP", None)
        for faithful in (True, False):
            r = b.aux_elliptic_pbr(faithful=faithful)
            print(rep, mode, "faithful" if faithful else "device", round(r["vertices_per_s"]), round(r["ms_per_timestep_median"], 1), round(r["ms_per_timestep_mean"], 1), flush=True)
