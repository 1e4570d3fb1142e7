"""The DSP fields of config::Settings (config.rs:85-98) and the three profiles
(default_settings.toml:108-140)."""
from dataclasses import dataclass

from . import _lib
from .err import raise_for


@dataclass
class Settings:
    work_rate: int = 12480
    resample_atten: float = 30.0
    resample_delta_freq: float = 1000.0
    resample_cutout: float = 4800.0
    demodulation_atten: float = 25.0
    wav_resample_atten: float = 40.0
    wav_resample_delta_freq: float = 0.1

    @staticmethod
    def profile(name):
        c = _lib.CSettings()
        raise_for(_lib.load().apt_profile_settings(name.encode(), c))
        wav = {"standard": (40.0, 0.1), "fast": (30.0, 0.2), "slow": (50.0, 0.05)}[name]
        return Settings(c.work_rate, c.resample_atten, c.resample_delta_freq, c.resample_cutout,
                        c.demodulation_atten, *wav)

    def to_c(self):
        return _lib.CSettings(int(self.work_rate), float(self.resample_atten), float(self.resample_delta_freq),
                              float(self.resample_cutout), float(self.demodulation_atten))
