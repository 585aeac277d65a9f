"""Driver for tests/test_run_reference.py — runs in a subprocess because importing the reference's top-level modules
(`train`, `inference`, `utils`, `models`, `dgl` ...) rewires sys.path / sys.modules.

  python tests/ref_loop_driver.py --ref DIR --work DIR --mode cpu    (no GPU: loops + dgl stand-in, oracle-backed model)
  python tests/ref_loop_driver.py --ref DIR --work DIR --mode gpu    (the engine behind the unmodified loops)

Prints one JSON object.  TEST INFRASTRUCTURE: the cpu mode plugs the CPU oracle in as `models.GraphGatedGCNModel`
so that the launcher's plumbing (stand-in `dgl`, dataset files, shims, hyper-parameter overrides) can be exercised
where no GPU exists; the product path never does that."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import run_reference as rr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--work", required=True)
    ap.add_argument("--mode", required=True, choices=["cpu", "gpu"])
    args = ap.parse_args()
    rr.install(args.ref)
    import models
    out = {}
    captured = {}

    if args.mode == "cpu":
        from oracle.gatedgcn_oracle import OracleModel

        class FakeModel(OracleModel):                       # reference call: model(g, x, e, pe)
            def forward(self, graph, x, e, pe):
                src, dst = graph.edges()
                s = super().forward(src.long(), dst.long(), graph.num_nodes(), e, pe)
                captured["scores"] = s.detach()
                return s

        models.GraphGatedGCNModel = FakeModel
        device = "cpu"
        data = rr.write_synthetic_dataset(os.path.join(args.work, "data"), genome_lens=(400_000, 300_000), leftover_len=25_000)
        hp = {"device": device, "num_epochs": 3, "dim_latent": 32, "num_gnn_layers": 2, "wandb_mode": "disabled",
              "batch_size_train": 1, "batch_size_eval": 1, "len_threshold": 20, "num_decoding_paths": 8}
        r = rr.run_train(data, args.work, hp, out="cpu")
        out["train"] = r
        mp = os.path.join(args.work, "m.pt")
        torch.save(torch.load(r["checkpoint"], map_location="cpu", weights_only=False)["model_state_dict"], mp)
        torch.manual_seed(0)
        ri, walks, contigs = rr.run_inference(data, mp, hp, device, reference_decode=True)
        out["inference"] = ri
    else:
        device = "cuda:0"
        torch.cuda.set_device(0)
        data = rr.write_synthetic_dataset(os.path.join(args.work, "data"))
        hp = {"device": device, "num_epochs": 3, "dim_latent": 128, "num_gnn_layers": 3, "wandb_mode": "disabled",
              "num_parts_metis_train": 116, "num_parts_metis_eval": 16, "len_threshold": 10, "num_decoding_paths": 16}
        out["full"] = rr.run_train(data, args.work, {**hp, "batch_size_train": 1, "batch_size_eval": 1}, out="full")
        out["minibatch"] = rr.run_train(data, args.work, {**hp, "batch_size_train": 4, "batch_size_eval": 4}, out="mb")
        mp = os.path.join(args.work, "m.pt")
        sd = torch.load(out["full"]["checkpoint"], map_location="cpu", weights_only=False)["model_state_dict"]
        torch.save(sd, mp)
        # capture what the reference's inference loop gets from model(g, x, e, pe)
        orig_forward = models.GraphGatedGCNModel.forward
        seen = []

        def forward(self, graph, x, e, pe):
            s = orig_forward(self, graph, x, e, pe)
            seen.append((graph, e.detach().clone(), pe.detach().clone(), s.detach().clone()))
            return s

        models.GraphGatedGCNModel.forward = forward
        ri, walks, contigs = rr.run_inference(data, mp, hp, device)
        models.GraphGatedGCNModel.forward = orig_forward
        out["inference"] = ri
        # the same graphs through the direct API (gnnome_assembly_b200.GraphGatedGCNModel on an AssemblyGraph built
        # from the raw edge list) and through the CPU oracle
        import gnnome_assembly_b200 as gg
        from oracle.gatedgcn_oracle import OracleModel, rel_err
        direct = gg.GraphGatedGCNModel(1, 2, 128, 16, 3, 64, True, 16)
        direct.load_state_dict(sd, strict=True)
        direct.eval().to(device)
        oracle = OracleModel(1, 2, 128, 16, 3, 64, True, 16)
        oracle.load_state_dict(sd, strict=True)
        errs_direct, errs_oracle = [], []
        for graph, e, pe, s in seen:
            src, dst = graph.edges()
            g2 = gg.AssemblyGraph(src.cpu().long(), dst.cpu().long(), graph.num_nodes())
            with torch.no_grad():
                s2 = direct(g2, None, e, pe)
                s3 = oracle(src.cpu().long(), dst.cpu().long(), graph.num_nodes(), e.cpu(), pe.cpu())
            errs_direct.append(float((s - s2).abs().max()))
            errs_oracle.append(rel_err(s, s3))
        out["logit_max_abs_diff_vs_direct_api"] = errs_direct
        out["logit_rel_err_vs_cpu_oracle"] = errs_oracle
        from gnnome_assembly_b200 import plan as gplan
        out["plan_stats"] = dict(gplan.PLAN_STATS)
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
