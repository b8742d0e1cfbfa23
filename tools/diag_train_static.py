"""diagnostic: static vs dynamic training path, per-stage forward differences and the worst gradient tensors"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_train import _small_train_setup, _grads, _rel, _pre_bn_biases
from pillarnet_lts_b200 import train

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
model, cfg, batches = _small_train_setup(5, prec)
ex = batches[0]
res = {}
probe = None
for mode in (False, True):
    train.set_static(mode)
    model.zero_grad(set_to_none=True)
    sp = model.reader(dict(points_batched=ex["points_batched"]))
    n = sp.table.count()
    feats = model.backbone(sp)
    bev = model.neck(feats)
    preds = model.bbox_head(bev)
    if probe is None:
        g = torch.Generator(device="cuda").manual_seed(9)
        probe = [{k: torch.randn(v.shape, device="cuda", generator=g) for k, v in p.items()} for p in preds]
    loss = sum((p[k].float() * w[k]).sum() for p, w in zip(preds, probe) for k in p) / 1000.0
    loss.backward()
    st = {"reader": sp.feat[:n].float().detach().clone()}
    for k, v in feats.items():
        st[k] = (v.feat[:v.table.count()] if hasattr(v, "feat") else v).float().detach().clone()
    for t, p in enumerate(preds):
        for k, v in p.items():
            st[f"t{t}.{k}"] = v.float().detach().clone()
    res[mode] = (st, _grads(model), float(loss))
train.set_static(False)
print("loss", res[False][2], res[True][2])
for k in res[False][0]:
    a, b = res[True][0][k], res[False][0][k]
    print("fwd %-12s rel %.3e  shape %s" % (k, _rel(a, b), tuple(a.shape)))
skip = _pre_bn_biases(model)
errs = sorted(((_rel(res[True][1][k], res[False][1][k]), k, res[False][1][k].numel()) for k in res[False][1] if k not in skip), reverse=True)
for e in errs[:25]:
    print("grad %.3e %s (%d)" % e)
print("median", np.median([e[0] for e in errs]))
