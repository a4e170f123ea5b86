"""BASELINE configs[4] once (50 000 x 2048, top-100): a few evaluations for `ncu` to attach to (tools/runs/r02_run16.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wealy_b200  # noqa: E402,F401
from wealy_b200 import evaluation as we  # noqa: E402
from wealy_b200.data import synth  # noqa: E402

s = synth.make_eval_set(50_000, 2048, seed=5, dist="lyric_covers_test", device="cuda", md5_ids=False)
plan = we.EvalPlan(s["c"], s["i"], s["c"], s["i"])
for _ in range(4):
    r = plan.run(s["z"], s["z"], topk=100)
torch.cuda.synchronize()
print("path", plan.last_topk_path(), plan.stage_ms())
