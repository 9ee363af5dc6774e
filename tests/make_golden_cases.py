"""Cases shared by tests/golden/make_golden.py (which runs the real reference) and the GPU tests."""

FDS_CASES = {   # name -> (event length, seed, tier, FilterDerivativeSegmenter kwargs)
    "default_thresholds": (6000, 61, "A", dict(low_threshold=1, high_threshold=2, cutoff_freq=1000., sampling_freq=1.e5)),
    "low_thresholds": (9000, 62, "B", dict(low_threshold=0.05, high_threshold=3, cutoff_freq=2000., sampling_freq=1.e5)),
    "fast_sampling": (12000, 63, "A", dict(low_threshold=0.02, high_threshold=0, cutoff_freq=2000., sampling_freq=2.5e5)),
}
