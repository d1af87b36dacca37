"""Regenerates tests/golden/checkpoints.json from the reference tree.

Run in the authoring container only (needs /root/reference, which does not
exist on the GPU box):   python tests/golden/make_golden.py

It decodes the BSON checkpoints the reference commits (written by
`@save "./checkpoint/mymodel.bson" p opt l_loss_train l_loss_val iter`,
case2/case2.jl:178) with a minimal BSON reader, and copies the printed weight
table of robertson/ReadMe.md:21-35 and the generating-mechanism constants of
the scripts.  No reference SOURCE is copied: only numeric fixtures.
"""
from __future__ import annotations

import json
import os
import struct
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "checkpoints.json")


def _cstring(b, i):
    j = b.index(b"\x00", i)
    return b[i:j].decode("utf8"), j + 1


def parse_doc(b, i=0, as_list=False):
    size = struct.unpack_from("<i", b, i)[0]
    end = i + size - 1
    i += 4
    out = [] if as_list else {}
    while i < end:
        t = b[i]; i += 1
        key, i = _cstring(b, i)
        if t == 0x01:
            v = struct.unpack_from("<d", b, i)[0]; i += 8
        elif t == 0x02:
            n = struct.unpack_from("<i", b, i)[0]; i += 4
            v = b[i:i + n - 1].decode("utf8"); i += n
        elif t == 0x03:
            v, i = parse_doc(b, i)
        elif t == 0x04:
            v, i = parse_doc(b, i, as_list=True)
        elif t == 0x05:
            n = struct.unpack_from("<i", b, i)[0]; i += 5
            v = bytes(b[i:i + n]); i += n
        elif t == 0x08:
            v = bool(b[i]); i += 1
        elif t == 0x0A:
            v = None
        elif t == 0x10:
            v = struct.unpack_from("<i", b, i)[0]; i += 4
        elif t == 0x12:
            v = struct.unpack_from("<q", b, i)[0]; i += 8
        else:
            raise ValueError(f"unhandled BSON type {t:#x}")
        if as_list:
            out.append(v)
        else:
            out[key] = v
    return out, end + 1


def resolve(doc, node):
    """Follow backrefs; decode Float64/Float32 arrays and scalars."""
    if isinstance(node, list):
        return [resolve(doc, x) for x in node]
    if isinstance(node, dict):
        tag = node.get("tag")
        if tag == "backref":
            return resolve(doc, doc["_backrefs"][node["ref"] - 1])
        if tag == "array":
            name = node["type"]["name"][-1] if isinstance(node.get("type"), dict) else None
            if name in ("Float64", "Float32") and isinstance(node.get("data"), bytes):
                fmt = "<%dd" if name == "Float64" else "<%df"
                w = 8 if name == "Float64" else 4
                return list(struct.unpack(fmt % (len(node["data"]) // w), node["data"]))
            if name == "Any" or name is None:
                return [resolve(doc, x) for x in node["data"]]
        if tag == "struct" and isinstance(node.get("type"), dict):
            name = node["type"]["name"][-1]
            if name == "Float32":
                return struct.unpack("<f", node["data"])[0]
            if name == "Float64":
                return struct.unpack("<d", node["data"])[0]
    return node


def load(path):
    with open(path, "rb") as f:
        b = f.read()
    doc, _ = parse_doc(b)
    return doc


def summarize(path, keys=("l_loss_train", "l_loss_val")):
    doc = load(path)
    out = {"p": resolve(doc, doc["p"]), "iter": doc.get("iter")}
    for k in keys:
        if k in doc:
            v = resolve(doc, doc[k])
            v = [float(x) for x in v if isinstance(x, (int, float))]
            if v:
                out[k] = {"first": v[0], "last": v[-1], "min": min(v), "n": len(v)}
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are committed, nothing to do")
    g = {
        "source": "decoded from the BSON checkpoints committed in DENG-MIT/CRNN (see make_golden.py)",
        "case2": summarize(f"{REF}/case2/checkpoint/mymodel.bson"),
        "robertson": summarize(f"{REF}/robertson/checkpoint/mymodel.bson"),
        "gene": summarize(f"{REF}/gene-regulatory-network/checkpoint/mymodel.bson"),
        # 164 CRNN + 130 MLP parameters (yeast-glycolysis/yeast_glycolysis.jl:112-114,137-145)
        "yeast": summarize(f"{REF}/yeast-glycolysis/checkpoint/mymodel.bson"),
        # generating mechanisms (case2/case2.jl:52-53, robertson/rober_crnn.jl:52, case1/case1.jl:27)
        "case2_true": {"logA": [18.60, 19.13, 7.93], "Ea": [14.54, 14.42, 6.47], "R": 1.98720425864083e-3},
        "robertson_true": {"k": [4e-2, 3e7, 1e4]},
        "case1_true": {"k": [0.1, 0.2, 0.13, 0.3]},
        # robertson/ReadMe.md:21-27,35: printed `hcat(w_in', w_b, w_out')` and slope of another trained model
        "robertson_readme": {
            "table": [
                [2.5, 1.61821, 1.82531, 16.4681, -1.31015e-5, -29199.4, -9.06493],
                [0.194654, 1.81441, 0.0, 24.4825, -5.07967, -2194.18, 5.16405],
                [0.0, 1.71672, 1.82568, 24.0435, 0.0916024, -16210.3, -0.0870566],
                [0.0, 0.0, 0.0, -13.1789, 0.0145153, 0.0873887, 0.0711276],
                [0.826669, 0.0, 0.0, 0.114739, -0.140722, 0.160659, 0.0953954],
                [1.68456, 0.0, 0.0, 7.43096, -1.36473e-6, 154.215, 1.44846e-9]],
            "slope": 1.0110600333418567},
    }
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT, {k: (len(v["p"]) if isinstance(v, dict) and "p" in v else None) for k, v in g.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
