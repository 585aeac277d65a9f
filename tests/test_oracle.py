"""CPU tests: the oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py), against its fp64 self and against the mailbox (dead-UDF) form."""
import os

import numpy as np
import pytest
import torch

from oracle.gatedgcn_oracle import OracleGatedGCN, OracleModel, bce_loss, grads_close, rel_err

CASES = ["ref_rand_d64_L2_bn", "ref_rand_d64_L2_ln", "ref_asm_d128_L3_bn"]


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, f"{name}.pt"), weights_only=False)


def _run_oracle(g, state_dict, dtype=torch.float32, grads=True):
    model = OracleModel(1, 2, g["d"], 16, g["L"], 64, g["batch_norm"], 16)
    model.load_state_dict(state_dict, strict=True)
    model = model.to(dtype)
    src = torch.from_numpy(g["src"].astype(np.int64))
    dst = torch.from_numpy(g["dst"].astype(np.int64))
    e = torch.from_numpy(g["e"]).to(dtype)
    pe = torch.from_numpy(g["pe"]).to(dtype)
    scores = model(src, dst, g["num_nodes"], e, pe)
    out = {"scores": scores.detach()}
    if grads:
        loss = bce_loss(scores, torch.from_numpy(g["y"]), g["pos_weight"])
        loss.backward()
        out["loss"] = float(loss.detach())
        out["grads"] = {k: p.grad.detach() for k, p in model.named_parameters()}
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_code(golden_dir, name):
    g = _load(golden_dir, name)
    out = _run_oracle(g, g["state_dict"])
    assert rel_err(out["scores"], g["scores"]) < 2e-6
    assert abs(out["loss"] - g["loss"]) < 1e-6 * max(1.0, abs(g["loss"]))
    assert grads_close(out["grads"], g["grads"], rtol=2e-4) == []


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp64_tiebreak(golden_dir, name):
    g = _load(golden_dir, name)
    o32 = _run_oracle(g, g["state_dict"], torch.float32, grads=False)
    o64 = _run_oracle(g, g["state_dict"], torch.float64, grads=False)
    assert rel_err(o32["scores"], o64["scores"]) < 1e-5


def test_oracle_checkpoint_golden(golden_dir, ckpt_path):
    g = _load(golden_dir, "ref_asm_ckpt15xchr19")
    sd = torch.load(ckpt_path, map_location="cpu")
    assert len(sd) == 266
    out = _run_oracle(g, sd, grads=False)
    assert rel_err(out["scores"], g["scores"]) < 5e-6


@pytest.mark.parametrize("batch_norm", [True, False])
def test_layer_matches_mailbox_form(batch_norm):
    torch.manual_seed(0)
    n, m, d = 40, 160, 16
    src = torch.randint(0, n - 5, (m,))
    dst = torch.randint(0, n - 5, (m,))
    layer = OracleGatedGCN(d, d, batch_norm).double()
    h = torch.randn(n, d, dtype=torch.float64)
    e = torch.randn(m, d, dtype=torch.float64)
    h1, e1 = layer(src, dst, n, h, e)
    h2, e2 = layer.forward_mailbox(src, dst, n, h, e)
    assert rel_err(h1, h2) < 1e-12 and rel_err(e1, e2) < 1e-12


def test_edge_permutation_invariance(golden_dir):
    g = _load(golden_dir, "ref_rand_d64_L2_bn")
    base = _run_oracle(g, g["state_dict"], grads=False)["scores"]
    perm = np.random.default_rng(0).permutation(len(g["src"]))
    g2 = dict(g, src=g["src"][perm], dst=g["dst"][perm], e=g["e"][perm], y=g["y"][perm])
    out = _run_oracle(g2, g["state_dict"], grads=False)["scores"]
    assert rel_err(out, base[torch.from_numpy(perm)]) < 1e-5
