"""Convert the reference's shipped synthetic datasets into compact .npz fixtures.

The reference ships dataset_/{density,cut_ratio,coreness,component}/tmp.npy as
pickled dicts holding a networkx graph (datasets.py:105-125).  /root/reference
does not exist on the GPU box, so the *data* (not code) is re-encoded here as
plain integer arrays: the raw edge list in networkx iteration order, the padded
subgraph node matrix and the integer labels.  glass_b200.datasets.load_dataset
rebuilds the BaseGraph from these exactly as datasets.py:103-126 does.

Run in the build container only:  python scripts/convert_shipped_datasets.py
"""
import os
import sys

import numpy as np

REF = os.environ.get("GLASS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ["density", "cut_ratio", "coreness", "component"]:
        obj = np.load(f"{REF}/dataset_/{name}/tmp.npy", allow_pickle=True).item()
        g = obj["G"]
        edge = np.array([[e[0] for e in g.edges], [e[1] for e in g.edges]], dtype=np.int32)
        nodes = [n for n in g.nodes]
        sub = obj["subG"]
        lmax = max(len(s) for s in sub)
        pad = -np.ones((len(sub), lmax), dtype=np.int32)
        for i, s in enumerate(sub):
            pad[i, :len(s)] = np.asarray(s, dtype=np.int32)
        label = np.array([ord(c) - ord("A") for c in obj["subGLabel"]], dtype=np.int8)
        np.savez_compressed(f"{OUT}/{name}.npz", edge=edge, n_node=np.int64(len(nodes)),
                            subG_pad=pad, label=label)
        print(name, "nodes", len(nodes), "edges", edge.shape, "subG", pad.shape,
              "classes", np.unique(label), os.path.getsize(f"{OUT}/{name}.npz"))


if __name__ == "__main__":
    sys.exit(main())
