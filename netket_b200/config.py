"""The configuration flags of the reference this path reads (netket/utils/config_flags.py): set them on this object or
through the environment variable of the same name in upper case, as in the reference.

    netket_experimental_fft_autocorrelation   NETKET_EXPERIMENTAL_FFT_AUTOCORRELATION=1
        `statistics` computes tau_corr from the full autocorrelation function with Sokal's window and adds `tau_corr_max`
        (netket/stats/mc_stats.py:296-331).
"""

import os


def _env_flag(name):
    return os.environ.get(name, "0").strip().lower() in ("1", "true", "yes", "on")


class _Config:
    def __init__(self):
        self.netket_experimental_fft_autocorrelation = _env_flag("NETKET_EXPERIMENTAL_FFT_AUTOCORRELATION")

    def update(self, name, value):
        """nk.config.update("netket_experimental_fft_autocorrelation", True)"""
        if not hasattr(self, name.lower()):
            raise AttributeError(f"unknown configuration flag {name!r}")
        setattr(self, name.lower(), value)


config = _Config()
