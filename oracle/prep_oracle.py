"""CPU oracle for the input-preparation and loss/metrics rows.  TEST INFRASTRUCTURE ONLY.
Each function restates the reference lines it cites; pinned to the reference's own utils.py by
tests/golden/ref_prep_small.pt (tests/golden/make_golden.py runs utils.preprocess_graph,
utils.add_positional_encoding and utils.calculate_tfpn on top of oracle/dgl_shim.py)."""
import numpy as np
import scipy.sparse as sp
import torch


def zscore_features(overlap_length, overlap_similarity):
    """utils.py:70-74."""
    ol_len = torch.as_tensor(overlap_length).float()
    ol_sim = torch.as_tensor(overlap_similarity).float()
    ol_len = (ol_len - ol_len.mean()) / ol_len.std()
    ol_sim = (ol_sim - ol_sim.mean()) / ol_sim.std()
    return torch.cat((ol_len.unsqueeze(-1), ol_sim.unsqueeze(-1)), dim=1)


def positional_encoding(src, dst, n, pe_dim=16, alpha=0.95):
    """utils.py:102-103,124-138 and the concat of train.py:249-251."""
    src, dst = np.asarray(src), np.asarray(dst)
    A = sp.csr_matrix((np.ones(len(src)), (src, dst)), shape=(n, n))   # g.adjacency_matrix: A[u, v] = #edges u->v
    D = A.sum(axis=1)
    Dinv = 1.0 / (D + 1e-9)
    Dinv[D < 1e-9] = 0
    Dinv = sp.diags(np.squeeze(np.asarray(Dinv)), dtype=float)
    P = (Dinv @ A).T
    x = np.ones([n]) / n
    cols = []
    for _ in range(pe_dim):
        x = alpha * P.dot(x) + (1.0 - alpha) / n * np.ones([n])
        cols.append(torch.from_numpy(x).float())
    pe = torch.stack(cols, dim=-1) if cols else torch.zeros(n, 0)
    in_deg = torch.from_numpy(np.bincount(dst, minlength=n)).float().unsqueeze(1)
    out_deg = torch.from_numpy(np.bincount(src, minlength=n)).float().unsqueeze(1)
    return torch.cat((in_deg, out_deg, pe), dim=1)


def bce_and_tfpn(scores, y, pos_weight):
    """train.py:210-211,255 and utils.py:217-223."""
    scores = scores.reshape(-1)
    pw = torch.tensor([pos_weight], dtype=scores.dtype)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=pw)(scores, y.to(scores.dtype))
    pred = torch.round(torch.sigmoid(scores))
    TP = torch.sum(torch.logical_and(pred == 1, y == 1)).item()
    TN = torch.sum(torch.logical_and(pred == 0, y == 0)).item()
    FP = torch.sum(torch.logical_and(pred == 1, y == 0)).item()
    FN = torch.sum(torch.logical_and(pred == 0, y == 1)).item()
    return loss, (TP, TN, FP, FN)
