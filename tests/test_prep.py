"""Input preparation (z-scored edge features, degree + PageRank positional encoding) and fused
BCE loss + TP/TN/FP/FN — SURVEY.md §8f rows 1-2.  The golden fixture was produced by the reference's own
utils.preprocess_graph / add_positional_encoding / calculate_tfpn (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import prep_oracle


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "ref_prep_small.pt"), weights_only=False)


# ------------------------------------------------------------------ CPU: oracle pinned to the reference's code
def test_oracle_zscore_matches_reference(gold):
    e = prep_oracle.zscore_features(torch.from_numpy(gold["overlap_length"].astype(np.int64)), gold["overlap_similarity"])
    assert torch.equal(e, gold["e"])


def test_oracle_pe_matches_reference(gold):
    pe = prep_oracle.positional_encoding(gold["src"], gold["dst"], gold["num_nodes"], 16)
    assert pe.shape == gold["pe"].shape
    assert torch.allclose(pe, gold["pe"], rtol=1e-6, atol=0)


def test_oracle_loss_and_tfpn_match_reference(gold):
    loss, tfpn = prep_oracle.bce_and_tfpn(gold["scores"], torch.from_numpy(gold["y"]), gold["pos_weight"])
    assert tfpn == tuple(gold["tfpn"])
    assert abs(float(loss) - gold["loss"]) < 1e-7


# ------------------------------------------------------------------ GPU
def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
def test_gpu_zscore_matches_reference(gold):
    dev = _dev()
    from gnnome_assembly_b200 import prep
    # the golden generator hands overlap_length to the reference as int64 (as Raven's integer overlaps are,
    # make_golden.py:78), which truncates the few fractional synthetic values: feed the same integers here
    ol = np.trunc(gold["overlap_length"]).astype(np.float32)
    e = prep.preprocess_features(torch.from_numpy(ol).to(dev), torch.from_numpy(gold["overlap_similarity"]).to(dev))
    # bar: the reference computes (x - mean) / std in fp32; its fp32 mean carries a few ulps of summation error
    # which x - mean turns into ~1e-5 absolute in the z-score (|z| ~ 1): noise of the reference, not signal.
    # The kernel subtracts and divides in fp64 and rounds once, so it is held (a) to the exact fp64 value within
    # fp32 rounding and (b) to the reference's values within that noise (2e-5 / 3e-4 absolute per column).
    ref64 = np.stack([(a - a.mean()) / a.std(ddof=1) for a in
                      (ol.astype(np.float64), gold["overlap_similarity"].astype(np.float64))], 1)
    assert np.abs(e.cpu().numpy() - ref64).max() < 5e-7
    assert torch.allclose(e.cpu()[:, 0], gold["e"][:, 0], rtol=0, atol=2e-5)
    assert torch.allclose(e.cpu()[:, 1], gold["e"][:, 1], rtol=0, atol=3e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("relabel", [True, False])
def test_gpu_pe_matches_reference(gold, relabel):
    dev = _dev()
    from gnnome_assembly_b200 import GraphPlan, prep
    plan = GraphPlan(torch.from_numpy(gold["src"]), torch.from_numpy(gold["dst"]), gold["num_nodes"], dev, relabel=relabel)
    pe = prep.positional_encoding(plan, 16)
    assert torch.equal(pe[:, :2].cpu(), gold["pe"][:, :2])                      # degrees: exact
    assert torch.allclose(pe.cpu(), gold["pe"], rtol=2e-6, atol=0)


@pytest.mark.gpu
def test_gpu_pe_full_size_vs_oracle():
    dev = _dev()
    from gnnome_assembly_b200 import GraphPlan, prep
    from gnnome_assembly_b200.synth import make_assembly_graph
    g = make_assembly_graph("chr19", seed=0)
    plan = GraphPlan(torch.from_numpy(g.src), torch.from_numpy(g.dst), g.num_nodes, dev)
    pe = prep.positional_encoding(plan, 16).cpu()
    ref = prep_oracle.positional_encoding(g.src, g.dst, g.num_nodes, 16)
    assert torch.allclose(pe, ref, rtol=2e-6, atol=0)
    assert torch.allclose(pe, torch.from_numpy(g.pe), rtol=2e-6, atol=0)         # the generator's own PE
    assert abs(float(pe[:, 2:].sum(0).max()) - 1.0) < 0.06                        # PageRank mass stays ~1


@pytest.mark.gpu
def test_gpu_loss_metrics_match_reference(gold):
    dev = _dev()
    from gnnome_assembly_b200 import prep
    s = gold["scores"].to(dev).requires_grad_()
    y = torch.from_numpy(gold["y"]).to(dev)
    loss, counts = prep.bce_with_logits_and_metrics(s, y, gold["pos_weight"])
    assert tuple(int(c) for c in counts.tolist()) == tuple(gold["tfpn"])
    assert abs(float(loss) - gold["loss"]) < 2e-7
    loss.backward()
    s2 = gold["scores"].clone().requires_grad_()
    l2, _ = prep_oracle.bce_and_tfpn(s2, torch.from_numpy(gold["y"]), gold["pos_weight"])
    l2.backward()
    assert torch.allclose(s.grad.cpu(), s2.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_gpu_loss_on_model_output_shape():
    """scores as the model returns them ([E,1]) and an empty graph."""
    dev = _dev()
    from gnnome_assembly_b200 import prep
    s = torch.randn(100, 1, device=dev, requires_grad=True)
    y = (torch.rand(100, device=dev) > 0.5).float()
    loss, counts = prep.bce_with_logits_and_metrics(s, y, 0.5)
    loss.backward()
    assert s.grad.shape == (100, 1) and int(counts.sum()) == 100
    loss0, c0 = prep.bce_with_logits_and_metrics(torch.zeros(0, 1, device=dev), torch.zeros(0, device=dev), 0.5)
    assert float(loss0) == 0.0 and int(c0.sum()) == 0
