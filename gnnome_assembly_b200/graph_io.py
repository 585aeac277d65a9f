"""DGL-free assembly-graph files — SURVEY.md §8f row 4 (second half).

The reference stores its graphs with `dgl.save_graphs(processed_path, graph)` (graph_dataset.py:129) and reads them
back with `dgl.load_graphs(path)[0][0]` (graph_dataset.py:72).  DGL's `.dgl` container is a DGL-internal binary that
cannot be read or written without libdgl, which is not installable here — so graphs produced without DGL (the
synthetic generator, or an export script run where DGL exists: `export_from_dgl`) use the container below.  The file
extension stays `.dgl` because the dataset derives the graph index from it (`int(file[:-4])`, graph_dataset.py:71).

Layout (little endian):
    bytes 0..7    magic  b"GGASMG01"
    bytes 8..15   uint64 header length H
    bytes 16..    H bytes of UTF-8 JSON:
                    {"graphs": [{"num_nodes": n, "num_edges": m, "idtype": "int32"|"int64",
                                 "src": [off, nbytes], "dst": [off, nbytes],
                                 "ndata": {name: {"dtype": "float32", "shape": [...], "span": [off, nbytes]}},
                                 "edata": {...}}],
                     "labels": {name: {"dtype", "shape", "span"}}}
    then          the raw arrays, each 64-byte aligned; offsets are relative to the first byte after the header.
Arrays are read with one `numpy.fromfile`-style slice each (memory-mapped), so a 20M-edge graph loads at disk speed.
"""
from __future__ import annotations

import json
import struct

import numpy as np
import torch

from .graph import AssemblyGraph

MAGIC = b"GGASMG01"
_ALIGN = 64


def _np(t):
    t = torch.as_tensor(t).detach().cpu().contiguous()
    return t.numpy()


class _Packer:
    def __init__(self):
        self.chunks, self.off = [], 0

    def add(self, arr):
        arr = np.ascontiguousarray(arr)
        pad = (-self.off) % _ALIGN
        if pad:
            self.chunks.append(b"\0" * pad)
            self.off += pad
        span = [self.off, arr.nbytes]
        self.chunks.append(arr.tobytes())
        self.off += arr.nbytes
        return {"dtype": str(arr.dtype), "shape": list(arr.shape), "span": span}


def save_graphs(path, g_list, labels=None):
    """dgl.save_graphs(filename, g_list, labels=None) (graph_dataset.py:129)."""
    if not isinstance(g_list, (list, tuple)):
        g_list = [g_list]
    pk = _Packer()
    graphs = []
    for g in g_list:
        src, dst = g.edges()
        s, d = _np(src), _np(dst)
        if s.dtype not in (np.int32, np.int64):
            s, d = s.astype(np.int64), d.astype(np.int64)
        graphs.append({
            "num_nodes": int(g.num_nodes()), "num_edges": int(s.shape[0]), "idtype": str(s.dtype),
            "src": pk.add(s)["span"], "dst": pk.add(d)["span"],
            "ndata": {k: pk.add(_np(v)) for k, v in g.ndata.items()},
            "edata": {k: pk.add(_np(v)) for k, v in g.edata.items()},
        })
    header = json.dumps({"graphs": graphs,
                         "labels": {k: pk.add(_np(v)) for k, v in (labels or {}).items()}}).encode()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(header)))
        f.write(header)
        for c in pk.chunks:
            f.write(c)


def load_graphs(path, idx_list=None):
    """dgl.load_graphs(filename) -> (graph list, label dict) (graph_dataset.py:72).  Graphs come back host-resident,
    like DGL's."""
    with open(path, "rb") as f:
        magic = f.read(8)
        if magic != MAGIC:
            raise ValueError(
                f"{path}: not a gnnome_assembly_b200 graph file (magic {magic!r}).  DGL's own .dgl binaries can only be "
                "read by DGL; convert them once where DGL is installed with gnnome_assembly_b200.graph_io.export_from_dgl")
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode())
        base = 16 + hlen
    blob = np.memmap(path, dtype=np.uint8, mode="r", offset=base) if _has_payload(header) else np.zeros(0, np.uint8)

    def arr(dtype, shape, span):
        off, nbytes = span
        a = np.frombuffer(blob[off:off + nbytes], dtype=np.dtype(dtype)).reshape(shape)
        return torch.from_numpy(a.copy())

    out = []
    for i, rec in enumerate(header["graphs"]):
        if idx_list is not None and i not in idx_list:
            continue
        m = rec["num_edges"]
        g = AssemblyGraph(arr(rec["idtype"], [m], rec["src"]), arr(rec["idtype"], [m], rec["dst"]), rec["num_nodes"])
        g.ndata = {k: arr(v["dtype"], v["shape"], v["span"]) for k, v in rec["ndata"].items()}
        g.edata = {k: arr(v["dtype"], v["shape"], v["span"]) for k, v in rec["edata"].items()}
        out.append(g)
    labels = {k: arr(v["dtype"], v["shape"], v["span"]) for k, v in header.get("labels", {}).items()}
    return out, labels


def _has_payload(header):
    for rec in header["graphs"]:
        if rec["num_edges"] or rec["ndata"] or rec["edata"]:
            return True
    return bool(header.get("labels"))


def export_from_dgl(dgl_path, out_path):
    """Convert a DGL `.dgl` binary to this container.  Needs a real DGL installation (run it where the reference's
    data was produced); kept here so that real Raven graphs can reach the engine."""
    import importlib
    dgl = importlib.import_module("dgl")
    if not hasattr(dgl, "__version__") or getattr(dgl, "__gnnome_standin__", False):
        raise RuntimeError("export_from_dgl needs the real DGL package, not the stand-in")
    graphs, labels = dgl.load_graphs(dgl_path)
    conv = []
    for g in graphs:
        s, d = g.edges()
        a = AssemblyGraph(s, d, g.num_nodes())
        a.ndata = {k: v for k, v in g.ndata.items()}
        a.edata = {k: v for k, v in g.edata.items()}
        conv.append(a)
    save_graphs(out_path, conv, labels)
