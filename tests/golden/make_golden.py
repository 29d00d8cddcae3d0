"""Generates the committed golden fixtures from the reference's own golden graph.

Run HERE (the build container), where /root/reference exists; the GPU box only sees the outputs:

    python tests/golden/make_golden.py

cnr2000_head.*  : nodes [0, 30000) of tests/data/cnr-2000 (arcs leaving the range dropped), recompressed
                  by the ORACLE's restatement of ANSBvGraph::store (window 7, max_ref 3, min_interval 4 --
                  the CLI defaults) and written in the reference's .ans/.pointers/.states layout.
cnr2000_head.npz: the expected CSR (what webgraph's BvGraphSeq yields for the same nodes) and the
                  (component, symbol) stream of the final BvComp pass.
cnr2000_full.json: whole-graph anchors (sizes, per-component model parameters, checksums).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import oracle_py as O  # noqa: E402
import wga_pkg  # noqa: E402

W = wga_pkg.load()
BASE = "/root/reference/tests/data/cnr-2000/cnr-2000"
HEAD = 30000


def main():
    off, succ, bits = O.read_bvgraph(BASE)
    # --- head sub-graph
    keep_off = [0]
    keep = []
    for v in range(HEAD):
        s = succ[off[v]:off[v + 1]]
        s = s[s < HEAD]
        keep.append(s)
        keep_off.append(keep_off[-1] + len(s))
    h_off = np.array(keep_off, np.uint64)
    h_succ = np.concatenate(keep).astype(np.uint32)
    g = O.OracleGraph.store_csr(h_off, h_succ, 7, 3, 4)
    inf = g.info()
    states, pointers = g.phases()
    W.write_files(os.path.join(HERE, "cnr2000_head"), g.tables(), g.stream(), inf["state"], inf["n"], 7, 4,
                  inf["arcs"], states, pointers)
    comps, syms = g.trace()
    np.savez_compressed(os.path.join(HERE, "cnr2000_head.npz"), offsets=h_off, succ=h_succ, comps=comps, syms=syms)
    # --- whole graph anchors
    anchors = {}
    for params in ((7, 3, 4), (7, 3, 2)):
        G = O.OracleGraph.store_csr(off, succ, *params)
        i = G.info()
        st, pt = G.phases()
        anchors["w%d_r%d_l%d" % params] = dict(
            stream_bytes=i["stream_len"] * 2, final_state=i["state"],
            models=[[t["frame_size"], t["fidelity"], t["radix"], len(t["entries"])] for t in G.tables()],
            stream_sha256=hashlib.sha256(G.stream().tobytes()).hexdigest(),
            states_sha256=hashlib.sha256(st.tobytes()).hexdigest(),
            pointers_sha256=hashlib.sha256(pt.tobytes()).hexdigest(),
            symbols=int(len(G.trace()[1])))
    anchors["graph"] = dict(nodes=int(len(off) - 1), arcs=int(len(succ)),
                            csr_sha256=hashlib.sha256(off.tobytes() + succ.tobytes()).hexdigest(),
                            ef_first=[int(x) for x in bits[:5]], ef_last=int(bits[-1]))
    with open(os.path.join(HERE, "cnr2000_full.json"), "w") as f:
        json.dump(anchors, f, indent=1, sort_keys=True)
    print("head graph:", inf, "files:", [(x, os.path.getsize(os.path.join(HERE, x))) for x in sorted(os.listdir(HERE))])


if __name__ == "__main__":
    main()
