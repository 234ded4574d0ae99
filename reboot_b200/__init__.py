"""reboot_b200 -- B200-native implementation of the ReBoot physics step (broadphase, sphere-sphere and
sphere-triangle narrowphase, resolution, per-body integration) behind the C ABI in
include/reboot_physics.h. CUDA only: there is no CPU fallback."""
from .world import World, RebootError  # noqa: F401
from ._lib import RB_ACTIVE, RB_GRAVITY, RB_CONTACT  # noqa: F401
from . import scenes  # noqa: F401

__version__ = "0.1.0"
