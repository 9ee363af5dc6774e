"""Mirror of PyPore/cparsers.pyx's public class: ``FastStatSplit`` with the reference's positional
constructor (cparsers.pyx:55-58), its public ``min_gain`` (cparsers.pyx:52) and ``parse`` /
``best_single_split`` / ``score_samples``, all running on the GPU through pypore_b200.parsers."""
from .parsers import SpeedyStatSplit


class FastStatSplit(SpeedyStatSplit):
    def __init__(self, min_width=100, max_width=1000000, window_width=10000, min_gain_per_sample=None,
                 false_positive_rate=None, prior_segments_per_second=None, sampling_freq=1.e5,
                 cutoff_freq=None):
        SpeedyStatSplit.__init__(self, min_width, max_width, window_width, min_gain_per_sample,
                                 false_positive_rate, prior_segments_per_second, sampling_freq, cutoff_freq)
        self._params()  # the reference validates in the constructor (cparsers.pyx:69-76)
