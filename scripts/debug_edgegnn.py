"""Debug: EdgeGNN golden, per-parameter gradient error, with ordered / atomic backward variants."""
import functools, os, sys
import numpy as np, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glass_b200 import models, ops
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DEV = "cuda:0"
def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu(); b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / max(1e-12, float(b.abs().max())))
def run(tag, emb_atomic=False, pool_atomic=False):
    d = np.load(os.path.join(GOLDEN, "model_edgegnn.npz"))
    H, L = 32, 2
    conv = models.EmbGConv(H, H, H, L, max_deg=11, activation=nn.ReLU(inplace=True), jk=True, dropout=0.0,
                           conv=functools.partial(models.MyGCNConv, aggr="mean", activation=nn.ReLU(inplace=True)), gn=True)
    m = models.EdgeGNN(conv, nn.ModuleList([nn.Linear(H * L, 1)]), nn.ModuleList([models.MeanPool()]))
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")}
    m.load_state_dict(sd); m = m.to(DEV).train()
    t = lambda k: torch.from_numpy(d[k]).to(DEV)
    orig_plan = ops._embed_plan
    orig_pool = ops._SegmentPool.backward
    if emb_atomic:
        ops._embed_plan = lambda ids: None
    if pool_atomic:
        def bwd(ctx, dout):
            pos, cnt, argmax = ctx.saved_tensors
            mode, n = ctx.cfg
            dout = dout.contiguous()
            demb = torch.zeros((n, dout.shape[1]), dtype=torch.float32, device=dout.device)
            torch.ops.glass_b200.segment_pool_bwd_(dout, pos, mode, cnt, argmax, demb, None)
            return demb, None, None
        ops._SegmentPool.backward = staticmethod(bwd)
    try:
        logits = m(t("x"), t("ei"), t("ew"), t("pairs"))
        loss = nn.BCEWithLogitsLoss()(logits.flatten(), t("y"))
        loss.backward()
    finally:
        ops._embed_plan = orig_plan
        ops._SegmentPool.backward = orig_pool
    print(tag, "logits", rel(logits.detach(), d["logits"]))
    for k, p in m.named_parameters():
        print(f"   {k:40s} {rel(p.grad, d['grad.' + k]):.3e}")
for i in range(2):
    run(f"ordered/ordered #{i}")
run("emb atomic", emb_atomic=True)
run("pool atomic", pool_atomic=True)
run("both atomic", True, True)
