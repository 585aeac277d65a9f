#!/usr/bin/env python
"""Run the reference's OWN, UNMODIFIED train.py / inference.py on the B200 engine.

    python tools/run_reference.py --ref /path/to/GNNome-assembly [--mode full|minibatch|inference|all]
                                  [--data DIR] [--work DIR] [--device cuda:0] [--set key=value ...]

north_star: "drops in behind the existing models/full_graph.py / layers/processor.py forward so train.py and
inference.py are unchanged".  The reference scripts are imported as they are; what this launcher does around them:

  1. sys.path order: <repo>/gnnome_assembly_b200/dropin first — `import models` / `import layers` resolve to the
     engine (same class names, ctor arguments, state_dict keys), and `import dgl` resolves to dropin/dgl, a graph
     holder + graph files + dataset base class + device-side cluster sampler (DGL itself cannot be installed here;
     with a real DGL, pass --real-dgl and only models / layers are swapped);
  2. three facts about this image that are not about the engine:
       * torch >= 2.4 removed ReduceLROnPlateau(verbose=) (train.py:212)  -> a subclass that accepts and drops it,
       * Biopython is absent (evaluate.py:6, graph_parser.py:4)          -> a 40-line `Bio.SeqIO` FASTA writer,
       * hyperparameters.py:25 hard-codes 'cuda:3'                        -> `get_hyperparameters` is wrapped so that
         --device / --set overrides are applied to the dict it returns (configuration, not code);
  3. inference.py's Python decoder (get_contigs, :182-259) is replaced by gnnome_assembly_b200.decode.get_contigs
     (same signature and result) unless --reference-decode is given;
  4. with no --data, a small synthetic dataset in the reference's on-disk layout (processed/<idx>.dgl in the
     DGL-free container of graph_io.py, info/<idx>_{succ,pred,edges,reads}.pkl, info/g_to_chr.pkl) is written first.

Prints one JSON line per phase with the per-graph losses parsed from the reference's own stdout.
"""
from __future__ import annotations

import argparse
import contextlib
import importlib
import io
import json
import os
import pickle
import re
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_REF = "/root/reference" if os.path.isdir("/root/reference") else os.path.join(ROOT, "baseline", "_ref")


# --------------------------------------------------------------------------------------------- shims
def _install_plateau_verbose_shim():
    import inspect
    import torch.optim.lr_scheduler as sched
    if "verbose" in inspect.signature(sched.ReduceLROnPlateau.__init__).parameters:
        return

    class ReduceLROnPlateau(sched.ReduceLROnPlateau):            # train.py:212 passes verbose=True
        def __init__(self, *args, verbose=False, **kwargs):
            super().__init__(*args, **kwargs)

    ReduceLROnPlateau.__module__ = sched.__name__
    sched.ReduceLROnPlateau = ReduceLROnPlateau


def _install_bio_standin():
    try:
        import Bio  # noqa: F401
        return
    except ImportError:
        pass
    bio, seqio, seq = types.ModuleType("Bio"), types.ModuleType("Bio.SeqIO"), types.ModuleType("Bio.Seq")

    class Seq(str):
        pass

    class SeqRecord:                                             # evaluate.py:44-46 sets id / description
        def __init__(self, seq, id="<unknown id>", description=""):
            self.seq, self.id, self.description = seq, id, description

        def __len__(self):
            return len(self.seq)

    def write(records, handle, fmt):                              # evaluate.py:55: SeqIO.write(contigs, path, 'fasta')
        if fmt != "fasta":
            raise NotImplementedError(fmt)
        n = 0
        with open(handle, "w") as f:
            for r in records:
                f.write(f">{r.id} {r.description}\n")
                s = str(r.seq)
                for i in range(0, len(s), 60):
                    f.write(s[i:i + 60] + "\n")
                n += 1
        return n

    def parse(handle, fmt):
        raise NotImplementedError("Bio stand-in: parsing reads is not on this path")

    seqio.SeqRecord, seqio.write, seqio.parse = SeqRecord, write, parse
    seq.Seq = Seq
    bio.SeqIO, bio.Seq = seqio, seq
    sys.modules.update({"Bio": bio, "Bio.SeqIO": seqio, "Bio.Seq": seq})


def install(ref_dir, real_dgl=False):
    """Prepare the interpreter so that `import train` / `import inference` pick up the reference's scripts with the
    engine's models / layers (and the dgl stand-in) behind them."""
    ref_dir = os.path.abspath(ref_dir)
    if not os.path.exists(os.path.join(ref_dir, "train.py")):
        raise FileNotFoundError(f"{ref_dir}: no train.py — pass --ref <GNNome-assembly checkout>")
    dropin = os.path.join(ROOT, "gnnome_assembly_b200", "dropin")
    for p in (ref_dir, ROOT):
        while p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, ROOT)
    if real_dgl:                                  # only models / layers are swapped: a directory holding just those two
        only = os.path.join(ROOT, "gnnome_assembly_b200", "dropin_models_only")
        os.makedirs(only, exist_ok=True)
        for name in ("models", "layers"):
            link = os.path.join(only, name)
            if not os.path.exists(link):
                os.symlink(os.path.join(dropin, name), link)
        sys.path.insert(0, only)
    else:
        sys.path.insert(0, dropin)
    sys.path.append(ref_dir)
    _install_plateau_verbose_shim()
    _install_bio_standin()
    for name in ("models", "layers", "dgl", "train", "inference", "utils", "graph_dataset", "hyperparameters",
                 "evaluate", "graph_parser", "algorithms"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith((ref_dir, dropin)):
            del sys.modules[name]
    import models                                   # noqa: F401  -> the engine
    assert models.GraphGatedGCNModel.__module__.startswith("gnnome_assembly_b200"), models.__file__


def patch_hyperparameters(module, overrides):
    """Wrap module.get_hyperparameters (imported from hyperparameters.py at train.py:22 / inference.py:14) so the
    returned dict carries the overrides.  The reference's file is not edited."""
    orig = module.get_hyperparameters
    base = getattr(orig, "__wrapped__", orig)

    def get_hyperparameters():
        hp = dict(base())
        hp.update(overrides)
        return hp

    get_hyperparameters.__wrapped__ = base
    module.get_hyperparameters = get_hyperparameters


# --------------------------------------------------------------------------------------------- synthetic dataset
def _union(parts):
    """Disjoint union of synthetic graphs; node counts are even, so the 2k / 2k+1 strand pairing survives the offsets."""
    import dataclasses
    names = ("y", "prefix_length", "overlap_length", "overlap_similarity", "read_length")
    off, cat = 0, {k: [] for k in ("src", "dst") + names}
    for p in parts:
        cat["src"].append(p.src.astype(np.int64) + off)
        cat["dst"].append(p.dst.astype(np.int64) + off)
        for k in names:
            cat[k].append(getattr(p, k))
        off += p.num_nodes
    return dataclasses.replace(parts[0], num_nodes=off, e=None, pe=None, **{k: np.concatenate(v) for k, v in cat.items()})


def write_synthetic_dataset(root, genome_lens=(1_500_000, 1_200_000), seed=0, with_reads=True, leftover_len=25_000):
    """A dataset directory in the layout AssemblyGraphDataset expects (graph_dataset.py:52-79): processed/<idx>.dgl,
    info/<idx>_{succ,pred,edges,reads}.pkl (graph_parser.py:12-73 builds these from graph.edges() in edge-id order),
    info/g_to_chr.pkl (inference.py:313).  Features as graph_parser.from_csv leaves them (raw overlap_length /
    overlap_similarity / prefix_length, read_length, labels y); preprocess_graph + add_positional_encoding run in the
    reference's own dataset code at load time."""
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from gnnome_assembly_b200.graph import AssemblyGraph
    from gnnome_assembly_b200.graph_io import save_graphs
    from gnnome_assembly_b200.synth import make_assembly_graph
    for d in ("raw", "raven_output", "processed", "info"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    rng = np.random.default_rng(seed)
    g_to_chr = {}
    for idx, gl in enumerate(genome_lens):
        # a chromosome-sized piece plus a short leftover component (real Raven graphs have them too).  The leftover
        # matters for the reference's own decoder: its loop (inference.py:192-247) only ends once the best sampled
        # walk is shorter than len_threshold, and on a graph with no edge left to sample Categorical raises.
        gs = _union([make_assembly_graph("chr19", seed=seed + idx, genome_len=gl, pe_dim=0),
                     make_assembly_graph("chr19", seed=seed + 100 + idx, genome_len=leftover_len, pe_dim=0)])
        g = AssemblyGraph(torch.from_numpy(gs.src.astype(np.int64)), torch.from_numpy(gs.dst.astype(np.int64)), gs.num_nodes)
        g.edata["overlap_length"] = torch.from_numpy(np.trunc(gs.overlap_length).astype(np.int64))
        g.edata["overlap_similarity"] = torch.from_numpy(gs.overlap_similarity.astype(np.float32))
        g.edata["prefix_length"] = torch.from_numpy(gs.prefix_length.astype(np.int64))
        g.edata["y"] = torch.from_numpy(gs.y.astype(np.float32))
        g.ndata["read_length"] = torch.from_numpy(gs.read_length.astype(np.int64))
        save_graphs(os.path.join(root, "processed", f"{idx}.dgl"), [g])
        succ = {i: [] for i in range(gs.num_nodes)}
        pred = {i: [] for i in range(gs.num_nodes)}
        edges = {}
        for i, (s, d) in enumerate(zip(gs.src.tolist(), gs.dst.tolist())):
            succ[s].append(d)
            pred[d].append(s)
            edges[(s, d)] = i
        info = os.path.join(root, "info")
        pickle.dump(succ, open(os.path.join(info, f"{idx}_succ.pkl"), "wb"))
        pickle.dump(pred, open(os.path.join(info, f"{idx}_pred.pkl"), "wb"))
        pickle.dump(edges, open(os.path.join(info, f"{idx}_edges.pkl"), "wb"))
        if with_reads:
            alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
            reads = {}
            for k in range(0, gs.num_nodes, 2):                  # node 2k / 2k+1: the two strands of one read
                seq = alphabet[rng.integers(0, 4, int(gs.read_length[k]))].tobytes().decode()
                reads[k] = seq
                reads[k + 1] = seq[::-1].translate(str.maketrans("ACGT", "TGCA"))
            pickle.dump(reads, open(os.path.join(info, f"{idx}_reads.pkl"), "wb"))
        g_to_chr[idx] = "chr19"
    pickle.dump(g_to_chr, open(os.path.join(root, "info", "g_to_chr.pkl"), "wb"))
    return root


# --------------------------------------------------------------------------------------------- phases
_LOSS_RE = re.compile(r"TRAINING \(one training graph\): Epoch = (\d+), Graph = (\d+)\s*\nLoss: ([0-9.eE+-]+|nan|inf)")
_VAL_RE = re.compile(r"VALIDATION \(one validation graph\): Epoch = (\d+), Graph = (\d+)\s*\nLoss: ([0-9.eE+-]+|nan|inf)")


class _Tee(io.TextIOBase):
    def __init__(self, echo):
        self.buf, self.echo = io.StringIO(), echo

    def write(self, s):
        self.buf.write(s)
        if self.echo:
            sys.__stdout__.write(s)
        return len(s)


def run_train(data_dir, work_dir, overrides, out="run", echo=False):
    """train.train(train_path, valid_path, out, overfit) of the reference (train.py:115) — full-graph branch when
    batch_size_train <= 1 (:243-259), cluster mini-batch branch otherwise (:282-312)."""
    train = importlib.import_module("train")
    patch_hyperparameters(train, overrides)
    os.makedirs(work_dir, exist_ok=True)
    cwd = os.getcwd()
    tee = _Tee(echo)
    t0 = time.perf_counter()
    try:
        os.chdir(work_dir)                            # train.py writes pretrained/ and checkpoints/ under the cwd
        with contextlib.redirect_stdout(tee):
            train.train(data_dir, data_dir, out, overfit=False)
    finally:
        os.chdir(cwd)
    text = tee.buf.getvalue()
    res = {"phase": "train", "seconds": time.perf_counter() - t0,
           "train_loss": [(int(a), int(b), float(c)) for a, b, c in _LOSS_RE.findall(text)],
           "valid_loss": [(int(a), int(b), float(c)) for a, b, c in _VAL_RE.findall(text)],
           "checkpoint": os.path.join(work_dir, "checkpoints", f"{out}.pt")}
    return res


def run_inference(data_dir, model_path, overrides, device, reference_decode=False, echo=False):
    """inference.inference(data_path, model_path, device) of the reference (inference.py:404)."""
    inference = importlib.import_module("inference")
    patch_hyperparameters(inference, overrides)
    if not reference_decode:
        from gnnome_assembly_b200 import decode
        inference.get_contigs = decode.get_contigs              # inference.py:487 looks the name up at call time
        inference.get_contigs_baselines = decode.get_contigs_baselines
    tee = _Tee(echo)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(tee):
        walks, contigs = inference.inference(data_dir, model_path, device)
    return {"phase": "inference", "seconds": time.perf_counter() - t0, "graphs": len(walks),
            "contigs_per_graph": [len(w) for w in walks], "walk_nodes_per_graph": [sum(len(c) for c in w) for w in walks],
            "contig_bases_per_graph": [sum(len(c.seq) for c in cs) for cs in contigs]}, walks, contigs


def _parse_overrides(items):
    out = {}
    for it in items or ():
        k, v = it.split("=", 1)
        try:
            out[k] = json.loads(v)
        except json.JSONDecodeError:
            out[k] = v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=DEFAULT_REF, help="GNNome-assembly checkout (unmodified)")
    ap.add_argument("--mode", default="all", choices=["full", "minibatch", "inference", "all"])
    ap.add_argument("--data", default=None, help="dataset directory in the reference's layout; default: synthetic")
    ap.add_argument("--work", default=None, help="working directory for pretrained/ and checkpoints/")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--set", action="append", default=[], help="hyper-parameter override key=value (JSON value)")
    ap.add_argument("--model", default=None, help="state_dict for --mode inference (default: the run's checkpoint)")
    ap.add_argument("--real-dgl", action="store_true", help="a real DGL is installed: swap only models / layers")
    ap.add_argument("--reference-decode", action="store_true", help="keep inference.py's Python get_contigs")
    ap.add_argument("--echo", action="store_true", help="echo the reference's own prints")
    args = ap.parse_args()

    import tempfile
    work = args.work or tempfile.mkdtemp(prefix="gg_ref_")
    install(args.ref, real_dgl=args.real_dgl)
    dev = torch.device(args.device)
    torch.cuda.set_device(dev)
    data = args.data or write_synthetic_dataset(os.path.join(work, "data"))
    # a few epochs of a mid-sized model by default; every key can be overridden with --set
    base = {"device": str(dev), "num_epochs": 2, "dim_latent": 128, "num_gnn_layers": 4,
            "num_parts_metis_train": 116, "num_parts_metis_eval": 16, "wandb_mode": "disabled"}
    base.update(_parse_overrides(args.set))
    ckpt = None
    if args.mode in ("full", "all"):
        r = run_train(data, work, {**base, "batch_size_train": 1, "batch_size_eval": 1}, out="full", echo=args.echo)
        ckpt = r["checkpoint"]
        print(json.dumps(r))
    if args.mode in ("minibatch", "all"):
        r = run_train(data, work, {**base, "batch_size_train": 4, "batch_size_eval": 4}, out="minibatch", echo=args.echo)
        ckpt = ckpt or r["checkpoint"]
        print(json.dumps(r))
    if args.mode in ("inference", "all"):
        model_path = args.model
        if model_path is None:
            if ckpt is None:
                raise SystemExit("--mode inference needs --model (a state_dict) when no training phase ran")
            model_path = os.path.join(work, "model_from_checkpoint.pt")
            torch.save(torch.load(ckpt, map_location="cpu", weights_only=False)["model_state_dict"], model_path)   # train.py:48-55: holds numpy scalars
        r, _, _ = run_inference(data, model_path, base, str(dev), args.reference_decode, echo=args.echo)
        print(json.dumps(r))


if __name__ == "__main__":
    main()
