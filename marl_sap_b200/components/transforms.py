"""Buffer preprocess transforms (mirror of /root/reference/src/components/transforms.py:4-22)."""
import torch as th


class Transform:
    def transform(self, tensor):
        raise NotImplementedError

    def infer_output_info(self, vshape_in, dtype_in):
        raise NotImplementedError


class OneHot(Transform):
    """actions [.., 1] -> one-hot [.., out_dim] (transforms.py:12-22).

    On a CUDA tensor the scatter is done by the `sap_onehot` kernel; the result is float32 like
    the reference's `.float()` and the caller casts it to the scheme dtype.
    """

    def __init__(self, out_dim):
        self.out_dim = out_dim

    def transform(self, tensor):
        if tensor.is_cuda:
            from .. import _lib

            lib = _lib.load()
            src = tensor.contiguous()
            out = th.empty(*tensor.shape[:-1], self.out_dim, dtype=th.float32, device=tensor.device)
            rows = src.numel()
            _lib.check(lib.sap_onehot(src.data_ptr(), _lib.sap_dtype(src.dtype), out.data_ptr(), _lib.SAP_F32, rows,
                                      self.out_dim, _lib.stream_ptr(tensor.device)), "sap_onehot")
            return out
        y_onehot = tensor.new_zeros(*tensor.shape[:-1], self.out_dim)
        y_onehot.scatter_(-1, tensor.long(), 1)
        return y_onehot.float()

    def infer_output_info(self, vshape_in, dtype_in):
        return (self.out_dim,), dtype_in
