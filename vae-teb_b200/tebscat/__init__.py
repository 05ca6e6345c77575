"""tebscat -- B200-native wavelet scattering for VAE-TEB (hot path only).

    from tebscat import Scattering1D                  # kymatio.torch.Scattering1D surface
    from tebscat import KymatioPhaseScattering1D      # hdf5_dataset/kymatio_phase_scattering.py surface
"""
from .torch_frontend import Scattering1D, ScatteringTorch1D
from .phase import KymatioPhaseScattering1D

__all__ = ['Scattering1D', 'ScatteringTorch1D', 'KymatioPhaseScattering1D']
__version__ = '0.1.0'
