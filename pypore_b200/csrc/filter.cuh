// filter.cuh -- K5: Event.filter (PyPore/DataTypes.py:258-274) = scipy.signal.filtfilt
// of a Bessel low-pass, as a parallel linear-recurrence scan.  (Kernels follow.)
#pragma once
#include "common.cuh"

constexpr int FILT_MAX_COEF = 9;  // filter order <= 8
