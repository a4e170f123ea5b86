python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import wealy_b200
from wealy_b200 import evaluation as we
from wealy_b200.data import synth
n=100000
s = synth.make_eval_set(n, 1024, seed=0, device="cuda", md5_ids=False)
z, c, i = s["z"], s["c"], s["i"]
aps, r1s = we.evaluate(c, i, z, c, i, z)
perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
aps_p, r1_p = we.evaluate(c, i, z, c[perm], i[perm], z[perm])
dr=(r1s-r1_p).abs(); da=(aps-aps_p).abs()
print("dr max",dr.max().item(),"frac",(dr>0).float().mean().item(),"n",(dr>0).sum().item())
print("da max",da.max().item(),"frac>1e-6",(da>1e-6).float().mean().item(),"map diff",(aps.double().mean()-aps_p.double().mean()).item())
idx=(dr>0).nonzero().flatten()[:10]
print(idx.tolist(), r1s[idx].tolist(), r1_p[idx].tolist())
import os
os.environ["WEALY_SYM"]="0"
aps_n, r1_n = we.evaluate(c, i, z, c, i, z)
dr=(r1_n-r1_p).abs(); print("nonsym vs perm: dr max",dr.max().item(),"n",(dr>0).sum().item())
dr=(r1_n-r1s).abs(); print("nonsym vs sym: dr max",dr.max().item(),"n",(dr>0).sum().item())
PY
