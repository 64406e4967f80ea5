"""Controller registry (/root/reference/src/controllers/__init__.py): only the shared-parameter MAC is on the
rollout hot path; jumpstart / non-shared / MADDPG controllers are out of scope."""
from .basic_controller import BasicMAC

REGISTRY = {"basic_mac": BasicMAC}
