"""vloam-noted_b200: B200-native lidar registration hot path (scanRegistration ->
laserOdometry -> laserMapping of liuzm-slam/VLOAM-NOTED) behind a C ABI."""
from . import _build  # noqa: F401
