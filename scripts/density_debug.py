import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from glass_b200 import datasets, utils, ops
from tests.helpers import GOLDEN, build_product_model
torch.cuda.set_device(0); DEV = "cuda:0"
d = np.load(os.path.join(GOLDEN, "trajectory_density.npz"))
params = json.loads(str(d["params"]))
ei, ew, n = datasets.load_edges("density")
raw = dict(H=params["hidden_dim"], L=params["conv_layer"], aggr=params["aggr"], z=params["z_ratio"], act="elu", jk=1, out=3, emb="one", pool=params["pool"])
m = build_product_model(raw, n)
m.load_state_dict({k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd.")})
m = m.to(DEV).train()
x = torch.ones((n, 1, 1), dtype=torch.int64, device=DEV)
pos = torch.from_numpy(d["pos"][0]).to(DEV)
y = torch.from_numpy(d["y"][0]).to(DEV)
res = {}
for path in ("simt", "auto"):
    ops.set_gemm_path(path)
    with torch.no_grad():
        emb = m.NodeEmb(x, ei.to(DEV), ew.to(DEV), utils.MaxZOZ(x, pos))
        out = m(x, ei.to(DEV), ew.to(DEV), pos, utils.MaxZOZ(x, pos))
    loss = torch.nn.CrossEntropyLoss()(out, y)
    res[path] = (emb, out, float(loss))
    print(path, "loss", float(loss), "ref", float(d["losses"][0]), "logits", out.flatten().tolist())
e0, e1 = res["simt"][0], res["auto"][0]
print("emb max abs diff", float((e0 - e1).abs().max()), "max", float(e0.abs().max()))
bad = ((e0 - e1).abs() > 1e-3).nonzero()
print("bad rows", torch.unique(bad[:, 0])[:20].tolist(), "count", bad.shape[0])
# layer-level check of the comb GEMM on the real operands
conv = m.conv.convs[0]
with torch.no_grad():
    mask = ops.label_mask(utils.MaxZOZ(x, pos))
    h0 = ops.embedding(x.reshape(-1), m.conv.input_emb.weight)
    h0 = m.conv.emb_gn(h0)
    from glass_b200._lib import GEMM_SIMT, GEMM_TCGEN05, ACT_NONE, ACT_ELU
    t0, t1 = conv.trans_fns
    xm = ops.pair_linear_mix(h0, None, t0.weight, t0.bias, t1.weight, t1.bias, mask, conv.z_ratio, ACT_ELU, GEMM_SIMT)
    yy = conv.gn(ops.spmm(conv.adj, xm))
    c0, c1 = conv.comb_fns
    a = ops.pair_linear_mix(yy, h0, c0.weight, c0.bias, c1.weight, c1.bias, mask, conv.z_ratio, ACT_NONE, GEMM_SIMT)
    b = ops.pair_linear_mix(yy, h0, c0.weight, c0.bias, c1.weight, c1.bias, mask, conv.z_ratio, ACT_NONE, GEMM_TCGEN05)
    print("comb simt vs tc: max abs diff", float((a - b).abs().max()), "scale", float(a.abs().max()), "h0 absmax", float(h0.abs().max()))
    badr = ((a - b).abs() > 1e-4 * a.abs().max()).nonzero()
    print("bad", badr.shape[0], badr[:10].tolist())
    if badr.shape[0]:
        r = int(badr[0, 0]); print("row", r, "simt", a[r].tolist(), "tc", b[r].tolist(), "yy", yy[r].tolist(), "h0", h0[r].tolist(), "mask", int(mask[r]))
