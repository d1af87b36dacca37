"""Reader / writer for the checkpoints the reference scripts keep with BSON.jl
(`@save "./checkpoint/mymodel.bson" p opt l_loss_train l_loss_val iter`, case2/case2.jl:178; `@load` at :183-186).

Host-side utility: in the Julia front end checkpointing stays Julia's own `@save` / `@load`; this module
lets the Python mirror of the training loop resume from a checkpoint the reference wrote and write one back in the same
dialect.  The wire format is plain BSON (bsonspec.org) carrying BSON.jl's lowering of Julia values:

  Vector{Float64}   {tag: "array", type: {tag: "datatype", params: [], name: ["Core", "Float64"]}, size: [n], data: <binary>}
  Vector{Any}       {tag: "array", type: {... name: ["Core", "Any"]}, size: [n], data: [ ... ]}
  Float32 scalar    {tag: "struct", type: {... name: ["Core", "Float32"]}, data: <4 bytes>}
  shared objects    {tag: "backref", ref: k} into the top-level `_backrefs` list (1-based)

`parse` / `encode` keep the integer widths and key order, so `encode(parse(b)) == b` holds byte for byte on the
reference's own files (tests/test_boundary_cpu.py, run where the reference tree is present).  Writing the `opt` object of a
fresh run (Flux optimiser structs with an IdDict keyed by `p`) is NOT attempted: `save` carries an `opt` subtree over
unchanged when it is given one that `load` returned, and omits it otherwise.  Julia cannot be run in this image: files
written here have not been loaded by BSON.jl itself.
"""
from __future__ import annotations

import struct

import numpy as np


class I32(int):
    """BSON int32 (0x10); plain Python ints are written as int64 (0x12)."""


class Bin(bytes):
    """BSON binary (0x05); `subtype` is the subtype byte."""
    subtype = 0

    def __new__(cls, data=b"", subtype=0):
        o = super().__new__(cls, data)
        o.subtype = subtype
        return o


def _cstring(b, i):
    j = b.index(b"\x00", i)
    return b[i:j].decode("utf8"), j + 1


def parse(b: bytes, i: int = 0, as_list: bool = False):
    """One BSON document starting at byte i -> (dict | list, next offset)."""
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
            v, i = parse(b, i)
        elif t == 0x04:
            v, i = parse(b, i, as_list=True)
        elif t == 0x05:
            n = struct.unpack_from("<i", b, i)[0]; sub = b[i + 4]; i += 5
            v = Bin(b[i:i + n], sub); i += n
        elif t == 0x08:
            v = bool(b[i]); i += 1
        elif t == 0x0A:
            v = None
        elif t == 0x10:
            v = I32(struct.unpack_from("<i", b, i)[0]); i += 4
        elif t == 0x12:
            v = struct.unpack_from("<q", b, i)[0]; i += 8
        else:
            raise ValueError(f"unhandled BSON element type {t:#x} at byte {i}")
        if as_list:
            out.append(v)
        else:
            out[key] = v
    return out, end + 1


def encode(doc) -> bytes:
    """dict | list -> BSON document bytes (inverse of `parse`)."""
    items = enumerate(doc) if isinstance(doc, list) else doc.items()
    body = bytearray()
    for k, v in items:
        key = str(k).encode("utf8") + b"\x00"
        if isinstance(v, bool):
            body += b"\x08" + key + (b"\x01" if v else b"\x00")
        elif isinstance(v, I32):
            body += b"\x10" + key + struct.pack("<i", int(v))
        elif isinstance(v, (int, np.integer)):
            body += b"\x12" + key + struct.pack("<q", int(v))
        elif isinstance(v, (float, np.floating)):
            body += b"\x01" + key + struct.pack("<d", float(v))
        elif isinstance(v, str):
            s = v.encode("utf8") + b"\x00"
            body += b"\x02" + key + struct.pack("<i", len(s)) + s
        elif isinstance(v, (bytes, bytearray)):
            body += b"\x05" + key + struct.pack("<i", len(v)) + bytes([getattr(v, "subtype", 0)]) + bytes(v)
        elif v is None:
            body += b"\x0A" + key
        elif isinstance(v, dict):
            body += b"\x03" + key + encode(v)
        elif isinstance(v, list):
            body += b"\x04" + key + encode(v)
        else:
            raise TypeError(f"cannot encode {type(v)} under key {k!r}")
    return struct.pack("<i", len(body) + 5) + bytes(body) + b"\x00"


# ---- BSON.jl's lowering of the Julia values the scripts save ----

def _datatype(name):
    return {"tag": "datatype", "params": [], "name": ["Core", name]}


def lower_array(a) -> dict:
    """numpy float64 / float32 vector or matrix -> BSON.jl array document (column-major payload, like Julia)."""
    a = np.asarray(a)
    name = {"float64": "Float64", "float32": "Float32", "int64": "Int64"}[a.dtype.name]
    return {"tag": "array", "type": _datatype(name), "size": [int(n) for n in a.shape],
            "data": Bin(np.asfortranarray(a).tobytes(order="F"))}


def resolve(doc: dict, node):
    """Follow backrefs and decode the numeric lowerings -> numpy arrays / Python floats; other nodes come back as parsed."""
    if isinstance(node, list):
        return [resolve(doc, x) for x in node]
    if isinstance(node, dict):
        tag = node.get("tag")
        if tag == "backref":
            return resolve(doc, doc["_backrefs"][node["ref"] - 1])
        t = node.get("type")
        name = t["name"][-1] if isinstance(t, dict) and isinstance(t.get("name"), list) else None
        if tag == "array":
            if name in ("Float64", "Float32", "Int64") and isinstance(node.get("data"), (bytes, bytearray)):
                dt = {"Float64": "<f8", "Float32": "<f4", "Int64": "<i8"}[name]
                return np.frombuffer(bytes(node["data"]), dtype=dt).reshape([int(n) for n in node["size"]], order="F").copy()
            if isinstance(node.get("data"), list):
                return [resolve(doc, x) for x in node["data"]]
        if tag == "struct" and name in ("Float32", "Float64") and isinstance(node.get("data"), (bytes, bytearray)):
            return float(struct.unpack("<f" if name == "Float32" else "<d", bytes(node["data"]))[0])
    return node


def load(path: str) -> dict:
    """-> dict(p [np] float64, iter, l_loss_train, l_loss_val (lists of floats, when present), opt (raw subtree or None),
    doc (the parsed document))."""
    with open(path, "rb") as f:
        b = f.read()
    doc, _ = parse(b)
    out = {"doc": doc, "p": np.asarray(resolve(doc, doc["p"]), dtype=np.float64).reshape(-1), "iter": doc.get("iter"),
           "opt": doc.get("opt"),
           "p_ref": doc["p"]["ref"] if isinstance(doc["p"], dict) and doc["p"].get("tag") == "backref" else None}
    for k, v in doc.items():
        if k.startswith("l_") or k.startswith("list_"):     # l_loss_train, l_loss_val, l_grad, list_loss_*, list_grad
            r = resolve(doc, v)
            out[k] = [float(x) for x in (r.tolist() if isinstance(r, np.ndarray) else r) if isinstance(x, (int, float))]
    return out


def lower_float32(x) -> dict:
    """Float32 scalar the way it sits inside the scripts' `Any[]` histories."""
    return {"tag": "struct", "type": _datatype("Float32"), "data": Bin(struct.pack("<f", float(x)))}


def save(path: str, p, iter_: int, opt=None, backrefs=None, p_ref=None, **lists):
    """Write `p`, `iter` and the histories (keyword arguments, e.g. l_loss_train=[...]) the way `@save` lowers them in the
    reference's own files: `p` a Vector{Float64}, `iter` an Int64, each history an `Any[]` of Float32 scalars.
    `opt` / `backrefs`: subtrees returned by `load` (doc["opt"], doc["_backrefs"]) to carry over unchanged; a fresh run
    writes no `opt` (see the module docstring)."""
    doc = {}
    if opt is not None:
        doc["opt"] = opt
    doc["iter"] = int(iter_)
    doc["p"] = lower_array(np.asarray(p, dtype=np.float64).reshape(-1))
    if opt is not None and backrefs is not None:
        # In the reference's files `p` is a backref to the SAME array object that keys the optimiser's IdDict state
        # (`{tag: backref, ref: k}`): keep that identity, or Flux would resume with fresh ADAM moments for the new `p`.
        # The carried-over `_backrefs` entry that held the old p receives the new values and `p` points at it again.
        # `p_ref`: the 1-based index `load` reports (out["p_ref"]) when the loaded file stored p that way.
        ref = p_ref
        if ref is not None:
            backrefs = list(backrefs)
            backrefs[ref - 1] = doc["p"]
            doc["p"] = {"tag": "backref", "ref": ref}
    for k, v in lists.items():
        doc[k] = [lower_float32(x) for x in v]
    if backrefs is not None:
        doc["_backrefs"] = backrefs
    with open(path, "wb") as f:
        f.write(encode(doc))
