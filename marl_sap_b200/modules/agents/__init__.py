"""Agent networks: plain torch modules, kept as in the reference (the only dense contraction on the path;
north star: "the agent network forward stays the reference's torch module").
/root/reference/src/modules/agents/rnn_agent.py:7-31, flat_const_agent.py:9-34."""
import torch as th
import torch.nn as nn
import torch.nn.functional as F


class _FcAgent(nn.Module):
    def __init__(self, input_shape, args, n_out):
        super().__init__()
        self.args = args
        self.fc1 = nn.Linear(input_shape, args.hidden_dim)
        if self.args.use_rnn:
            self.rnn = nn.GRUCell(args.hidden_dim, args.hidden_dim)
        else:
            self.rnn = nn.Linear(args.hidden_dim, args.hidden_dim)
        self.fc2 = nn.Linear(args.hidden_dim, n_out)

    def init_hidden(self):
        return self.fc1.weight.new(1, self.args.hidden_dim).zero_()

    def _wt(self, layer):
        """Cached contiguous W^T of a layer (refreshed when the parameter is updated in place or moved)."""
        cache = self.__dict__.setdefault("_wt_cache", {})
        key = id(layer)
        ver = layer.weight._version
        hit = cache.get(key)
        if hit is None or hit[0] != ver or hit[1].device != layer.weight.device:
            hit = (ver, layer.weight.detach().t().contiguous())
            cache[key] = hit
        return hit[1]

    def _linear(self, layer, x, relu=False):
        """``relu?(layer(x))`` for the rollout (CUDA, no grad).  Same fp32 math as ``F.relu(F.linear(x, W, b))`` - the
        contraction is still a torch/cuBLAS sgemm, equal to F.linear up to the accumulation order cuBLAS picks per shape
        (bit-identical at the bench shapes; tests hold it to 1e-5).  Only the way the bias / ReLU epilogue is issued
        differs, because cuBLAS's own epilogues are slow for these skinny fp32 GEMMs (B200, 409 600 rows, three layers:
        F.linear + F.relu 1.32 ms; addmm with a cached W^T 0.95 ms; this 0.80 ms):

        * wide input (fc1, K = 490): plain ``mm`` + the in-place bias/ReLU kernel ``sap_bias_act`` (0.47 + 0.03 ms; the
          beta*C epilogue of ``addmm`` costs 0.12 ms here);
        * narrow input (K = hidden): ``addmm`` with the ReLU fused by ``torch._addmm_activation`` when available."""
        if not (x.is_cuda and x.dim() == 2 and layer.bias is not None and x.dtype == th.float32) or th.is_grad_enabled():
            y = layer(x)
            return F.relu(y) if relu else y
        wt = self._wt(layer)
        if x.shape[1] >= 256 and x.is_contiguous():
            from ... import _lib

            y = th.mm(x, wt)
            self.__dict__["kernel_launches"] = self.__dict__.get("kernel_launches", 0) + 1  # sap_bias_act launches so far
            _lib.check(_lib.load().sap_bias_act(y.data_ptr(), layer.bias.data_ptr(), y.shape[0], y.shape[1], int(relu),
                                                _lib.stream_ptr(x.device)), "sap_bias_act")
            return y
        if relu and hasattr(th, "_addmm_activation"):
            return th._addmm_activation(layer.bias, x, wt)
        y = th.addmm(layer.bias, x, wt)
        return y.relu_() if relu else y

    # ------------------------------------------------------------------ opt-in: split-precision first layer
    SPLIT_SCALE = 2.0 ** -11

    def _split_weight(self, layer, terms, k_pad):
        """fc1's fp32 weight as `terms` fp16 pieces side by side, W = W_0 + 2^-11 W_1 (+ 2^-22 W_2), transposed and
        zero-padded to `k_pad` input rows: [k_pad, terms * hidden] fp16.  Each residual is scaled back by 2^11 before it is
        rounded, so every piece sits in fp16's normal range; 3 pieces carry 33 mantissa bits (>= fp32's 24)."""
        cache = self.__dict__.setdefault("_split_cache", {})
        key = (id(layer), terms, k_pad)
        ver = layer.weight._version
        hit = cache.get(key)
        if hit is None or hit[0] != ver or hit[1].device != layer.weight.device:
            r = layer.weight.detach().float()
            pieces = []
            for _ in range(terms):
                piece = r.half()
                pieces.append(piece)
                r = (r - piece.float()) * (1.0 / self.SPLIT_SCALE)
            cat = th.cat([x.t() for x in pieces], dim=1)  # [K, terms * H]
            w = cat.new_zeros(k_pad, cat.shape[1])
            w[:cat.shape[0]] = cat
            hit = (ver, w.contiguous())
            cache[key] = hit
        return hit[1]

    def _linear_fp16_split(self, layer, x16, relu):
        """``relu?(layer(x))`` for fp16 inputs (the observation rows exactly as the buffer holds them; the default path
        widens them to fp32 first, basic_controller.py:82): ONE fp16 tensor-core GEMM against [W_0 | W_1 | W_2] with fp32
        accumulation, then ``sap_split_bias_act`` folds the pieces, adds the bias and applies the ReLU.  Every product
        x * W_k is exact in fp32, so the result differs from the fp32 sgemm only in accumulation order (tests: <= 2e-6
        relative).  ``x16`` may carry zero pad columns (row pitch a multiple of 8: what the tensor-core kernels need)."""
        from ... import _lib

        terms = int(getattr(self.args, "agent_fc1_terms", 3))
        w = self._split_weight(layer, terms, x16.shape[1])
        y = th.mm(x16, w, out_dtype=th.float32)
        hidden = layer.weight.shape[0]
        out = th.empty(x16.shape[0], hidden, dtype=th.float32, device=x16.device)
        self.__dict__["kernel_launches"] = self.__dict__.get("kernel_launches", 0) + 1
        _lib.check(_lib.load().sap_split_bias_act(y.data_ptr(), terms, self.SPLIT_SCALE, layer.bias.data_ptr(), out.data_ptr(),
                                                  y.shape[0], hidden, int(relu), _lib.stream_ptr(x16.device)),
                   "sap_split_bias_act")
        return out

    def forward(self, inputs, hidden_state):
        if inputs.dtype == th.float16:
            if th.is_grad_enabled() or not inputs.is_cuda:
                raise RuntimeError("fp16 agent inputs are a rollout-only path (CUDA, torch.no_grad())")
            x = self._linear_fp16_split(self.fc1, inputs, relu=True)
        else:
            if inputs.shape[-1] != self.fc1.weight.shape[1]:  # padded staging rows of the fp16 path read at fp32
                inputs = inputs[..., :self.fc1.weight.shape[1]]
            x = self._linear(self.fc1, inputs, relu=True)
        h_in = hidden_state.reshape(-1, self.args.hidden_dim)
        if self.args.use_rnn:
            h = self.rnn(x, h_in)
        else:
            h = self._linear(self.rnn, x, relu=True)
        q = self._linear(self.fc2, h)
        return q, h


class RNNAgent(_FcAgent):
    """fc1 -> (GRUCell | Linear+ReLU) -> fc2 -> m Q-values (rnn_agent.py:7-31)."""

    def __init__(self, input_shape, args):
        super().__init__(input_shape, args, args.m)


class FlatConstellationAgent(_FcAgent):
    """Same trunk, M+1 outputs: Q for the top-M tasks + an "anything else" baseline (flat_const_agent.py:9-34)."""

    def __init__(self, input_shape, args):
        self.M = args.env_args["M"]
        super().__init__(input_shape, args, self.M + 1)


REGISTRY = {"rnn": RNNAgent, "flat_const_agent": FlatConstellationAgent}
