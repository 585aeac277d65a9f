"""Build the read-length quantile tables the synthetic assembly-graph generator samples from.

Run in the build container only (needs /root/reference):  python tools/make_length_quantiles.py
Source: data/references/lengths/chr{19,21}.txt of the reference (one HiFi read length per line,
consumed by seqrequester `-distribution`, pipeline.py:167-168).  We keep a 4097-point inverse CDF
per chromosome (32 KB) instead of the 110k-line files, so the generator works on the GPU box.
"""
import numpy as np

REF = "/root/reference/data/references/lengths"
out = {}
for chrom in ("chr19", "chr21"):
    lens = np.loadtxt(f"{REF}/{chrom}.txt", dtype=np.int64)
    q = np.quantile(lens, np.linspace(0.0, 1.0, 4097)).astype(np.float64)
    out[chrom] = q
    print(chrom, len(lens), "mean", lens.mean(), "table mean", q.mean())
np.savez_compressed("gnnome_assembly_b200/data/read_length_quantiles.npz", **out)
