"""Episode-buffer preprocess transforms.

Interface of /root/reference/src/components/transforms.py:4-22 (``transform`` + ``infer_output_info``); only
``OneHot`` exists there.  On CUDA tensors the expansion is done by the ``sap_onehot`` kernel.
"""
import torch as th


class Transform:
    """A transform maps a stored field to a derived field and tells the buffer the derived shape / dtype."""

    def transform(self, tensor):
        raise NotImplementedError(f"{type(self).__name__}.transform")

    def infer_output_info(self, vshape_in, dtype_in):
        raise NotImplementedError(f"{type(self).__name__}.infer_output_info")


class OneHot(Transform):
    """Integer class ids ``[..., 1]`` -> indicator vectors ``[..., out_dim]`` (float32, cast later by the buffer)."""

    def __init__(self, out_dim):
        self.out_dim = int(out_dim)

    def infer_output_info(self, vshape_in, dtype_in):
        # the reference keeps the ids' dtype for the derived field (transforms.py:21-22)
        return (self.out_dim,), dtype_in

    def transform(self, tensor):
        lead = tuple(tensor.shape[:-1])
        if tensor.is_cuda:
            from .. import _lib

            ids = tensor.contiguous()
            out = th.empty(lead + (self.out_dim,), dtype=th.float32, device=ids.device)
            rc = _lib.load().sap_onehot(ids.data_ptr(), _lib.sap_dtype(ids.dtype), out.data_ptr(), _lib.SAP_F32,
                                        ids.numel(), self.out_dim, _lib.stream_ptr(ids.device))
            _lib.check(rc, "sap_onehot")
            return out
        classes = th.arange(self.out_dim, dtype=th.int64).view((1,) * len(lead) + (self.out_dim,))
        return (tensor.to(th.int64) == classes).to(th.float32)
