"""One chunked all-vs-all evaluation (12 500 tracks x 8 chunks x 1024, redux=min) a few times: the command ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wealy_b200  # noqa: E402,F401
from wealy_b200 import evaluation as we  # noqa: E402
from wealy_b200.data import synth  # noqa: E402

n, ch, d = 12_500, 8, 1024
base = synth.make_eval_set(n, d, seed=8, device="cuda", md5_ids=False)
g = torch.Generator(device="cuda").manual_seed(108)
z = (base["z"][:, None, :] + 0.8 * base["z"].norm(dim=1).mean() / d ** 0.5 * torch.randn(n, ch, d, generator=g, device="cuda")).contiguous()
plan = we.EvalPlan(base["c"], base["i"], base["c"], base["i"])
for _ in range(4):
    res = plan.run(z, z, redux="min", allow_empty=True)
torch.cuda.synchronize()
print("map", float(res["sums"][0] / res["sums"][2]), "sweep ms", plan.last_sweep_ms())
