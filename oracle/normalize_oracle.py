"""TEST INFRASTRUCTURE ONLY -- numpy restatement of what the reference does to a stored scattering record
(`fhr_st`) between the HDF5 file and the model: trim, normalize_tensor_data, transposition.

Follows hdf5_dataset/hdf5_dataset.py (CombinedHDF5Dataset.__getitem__ :733-759, normalize_tensor_data :18-137).
Pinned against the live reference function by oracle/make_golden.py (tests/golden/normalize_fhr_st.npz).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np


def log_channels_of(config, n_channels):
    """:86-91 -- 'all_except_0' or an explicit list."""
    if config == 'all_except_0':
        return [c for c in range(n_channels) if c != 0]
    return list(config) if isinstance(config, (list, tuple)) else []


def asinh_channels_of(config, n_channels):
    """:94-99 -- 'all' or an explicit list."""
    if config == 'all':
        return list(range(n_channels))
    return list(config) if isinstance(config, (list, tuple)) else []


def normalize_record(S, mean, variance, log_channels=(), asinh_channels=(), log_epsilon=1e-6, trim=0,
                     time_major=True, dtype=np.float32):
    """S: (..., C, L) scattering coefficients.  Returns (..., L - 2 trim, C) if time_major else (..., C, L - 2 trim).

    trim (:733-741)  ->  log(clamp(x, min=0) + eps) on log_channels (:103-112)  ->  asinh on asinh_channels
    (:115-124)  ->  (x - mean) / (std + 1e-8) with std = sqrt(variance) in float32 (:62-66, :133-135)  ->
    transpose (:758-759)."""
    x = np.array(S, dtype=dtype)
    if trim > 0:
        x = x[..., trim:-trim]
    mean_t = np.asarray(mean, np.float32).astype(dtype)[:, None]
    std_t = np.asarray(np.sqrt(variance), np.float32).astype(dtype)[:, None]
    lc = list(log_channels)
    if lc:
        x[..., lc, :] = np.log(np.maximum(x[..., lc, :], dtype(0.0)) + dtype(log_epsilon))
    ac = list(asinh_channels)
    if ac:
        x[..., ac, :] = np.arcsinh(x[..., ac, :])
    x = (x - mean_t) / (std_t + dtype(1e-8))
    return np.swapaxes(x, -1, -2) if time_major else x
