import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import mcp_b200, riskcases
ctx = mcp_b200.Context(0)
E, S = riskcases.synthetic(6, 500, 8, 5)
E[0] = 0.0
E[1] = -E[2]
S[100] = S[200]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which != "all":
    E = E[int(which):int(which) + 1]
ctx.upload_exposures(E); ctx.upload_scenarios(S)
ctx.run(0.97)
print("ok", ctx.fetch()["var_trial"])
