"""Runner registry with the reference's keys (/root/reference/src/runners/__init__.py:1-7)."""
from .vec_runner import CudaVecRunner, EpisodeRunner, ParallelRunner

REGISTRY = {}
REGISTRY["episode"] = EpisodeRunner
REGISTRY["parallel"] = ParallelRunner
