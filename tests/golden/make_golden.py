"""Generates tests/golden/*.json|npz from the REFERENCE'S OWN sources compiled in place (oracle/_ref):
    src/hashing/syncmer.cpp, src/hixf/build/hierarchical_interleaved_xor_filter.hpp (DFS), src/hixf/search/*.
Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The vectors are small and committed; the GPU box never reads /root/reference.  Third-party pieces that are absent
from the reference tree (wyhash, IXF probe) enter through the labelled stand-ins of oracle/stubs and stay
"parity unpinned"; everything else below is pinned by the reference's code itself.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.oracle import HixfArrays, Reference  # noqa: E402
from taxor_b200 import tools  # noqa: E402


def main():
    ref = Reference()
    rng = np.random.default_rng(20260101)
    # ---- syncmer scan (tie rules included) ----
    seqs = {
        "random_300": rng.integers(0, 4, 300),
        "random_2k": rng.integers(0, 4, 2000),
        "polyA_120": np.zeros(120, int),
        "polyT_90": np.full(90, 3),
        "AT_repeat_200": np.resize([0, 3], 200),
        "ACG_repeat_150": np.resize([0, 1, 2], 150),
        "period5_260": np.resize(rng.integers(0, 4, 5), 260),
        "period11_400": np.resize(rng.integers(0, 4, 11), 400),
        "AT_only_500": rng.integers(0, 2, 500) * 3,
        "short_21": rng.integers(0, 4, 21),
        "exact_22": rng.integers(0, 4, 22),
        "with_N": np.concatenate([rng.integers(0, 4, 80), [4, 4], rng.integers(0, 4, 90)]),
    }
    half = rng.integers(0, 4, 60)
    seqs["rc_palindrome_120"] = np.concatenate([half, (3 - half)[::-1]])
    mix = rng.integers(0, 4, 900)
    mix[200:260] = 0
    mix[500:580] = np.resize([2, 1], 80)
    seqs["random_with_repeats_900"] = mix
    sync = []
    for name, codes in seqs.items():
        codes = np.asarray(codes, dtype=np.uint8)
        for (k, s, t) in [(22, 12, 5), (20, 10, 5), (16, 8, 4), (22, 12, 1), (22, 12, 11), (21, 11, 5)]:
            h = ref.syncmer_hashes(codes, k, s, t)
            sync.append(dict(name=name, seq="".join("ACGTN"[c] for c in codes), k=k, s=s, t=t, hashes=[str(int(x)) for x in h]))
    with open(os.path.join(HERE, "syncmer_golden.json"), "w") as f:
        json.dump(sync, f)

    # ---- thresholds ----
    thr = []
    counts = [0, 1, 2, 10, 100, 256, 257, 500, 873, 907, 1000, 2048, 4999, 10000, 49981, 123456, 1000000]
    for (w, k, p, e, syn) in [(20, 22, -1.0, 0.05, 1), (20, 22, -1.0, 0.04, 1), (20, 22, -1.0, 0.15, 1), (20, 22, -1.0, 0.045, 1),
                              (20, 20, -1.0, 0.0, 1), (20, 30, -1.0, 0.2, 1), (20, 12, -1.0, 0.1, 1),
                              (20, 20, -1.0, 0.05, 0), (20, 20, -1.0, 0.1, 0), (22, 22, -1.0, 0.02, 0), (24, 20, -1.0, 0.05, 0),
                              (20, 22, 0.3, 0.05, 1), (20, 20, 0.75, 0.05, 0), (20, 22, 1.0, 0.05, 1)]:
        t = ref.thresholder(w, k, p, e, syn)
        for c in counts:
            for sf in (1.0, 0.25):
                if c == 0 and not syn and p <= 0:
                    continue  # 0 k-mers: sqrt of a negative variance under -Ofast, undefined in the reference
                thr.append(dict(window=w, k=k, percentage=p, error_rate=e, use_syncmer=syn, count=c, scaling_factor=sf,
                                threshold=str(ref.threshold_get(t, c, sf))))
    with open(os.path.join(HERE, "threshold_golden.json"), "w") as f:
        json.dump(thr, f)

    # ---- HIXF DFS: a small 3-level hierarchy with split and merged bins ----
    ub = [np.unique(rng.integers(0, 2**63, size=int(n), dtype=np.uint64)) for n in rng.integers(150, 900, 40)]
    hx = tools.BuiltHixf(ub, t_max=4, seed=11)
    arrays = HixfArrays(hx.seed, hx.bins, hx.tbins, hx.seg_len, hx.data, hx.bin_off, hx.next_ixf_id, hx.bin_to_ub)
    rh = ref.make_hixf(arrays)
    queries = []
    for i in range(24):
        src = int(rng.integers(0, len(ub)))
        vals = np.concatenate([rng.choice(ub[src], size=int(rng.integers(1, len(ub[src]))), replace=False),
                               rng.integers(0, 2**63, size=int(rng.integers(0, 300)), dtype=np.uint64)])
        for thr_v in (0, 1, len(vals) // 4, len(vals) // 2, len(vals), len(vals) + 1, 2**40):
            u, c = ref.bulk_contains(rh, vals, int(thr_v))
            queries.append(dict(values=vals, threshold=int(thr_v), ub=u, cnt=c))
    ref.free_hixf(rh)
    np.savez_compressed(os.path.join(HERE, "dfs_golden.npz"),
                        seed=hx.seed, bins=hx.bins, tbins=hx.tbins, seg_len=hx.seg_len, bin_off=hx.bin_off,
                        next_ixf_id=hx.next_ixf_id, bin_to_ub=hx.bin_to_ub, data=np.concatenate(hx.data),
                        n_queries=len(queries),
                        **{f"q{i}_values": q["values"] for i, q in enumerate(queries)},
                        **{f"q{i}_threshold": np.array([q["threshold"]], dtype=np.uint64) for i, q in enumerate(queries)},
                        **{f"q{i}_ub": q["ub"] for i, q in enumerate(queries)},
                        **{f"q{i}_cnt": q["cnt"] for i, q in enumerate(queries)})
    print("golden vectors written:", len(sync), "syncmer cases,", len(thr), "threshold cases,", len(queries), "DFS queries")


if __name__ == "__main__":
    main()
