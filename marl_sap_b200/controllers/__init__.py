"""Controller registry (/root/reference/src/controllers/__init__.py): the shared-parameter MAC of the rollout hot path
and its jump-start variant; non-shared / MADDPG controllers are out of scope."""
from .basic_controller import BasicMAC
from .jumpstart_controller import JumpstartMAC

REGISTRY = {"basic_mac": BasicMAC, "jumpstart_mac": JumpstartMAC}
