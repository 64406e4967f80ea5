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

    @staticmethod
    def _linear(layer, x):
        """layer(x) as one cuBLAS call with the bias as beta*C (same fp32 math as F.linear, bit-identical results,
        without the separate bias-epilogue kernel cublasLt launches for these skinny fp32 GEMMs)."""
        if x.is_cuda and x.dim() == 2 and layer.bias is not None:
            return th.addmm(layer.bias, x, layer.weight.t())
        return layer(x)

    def forward(self, inputs, hidden_state):
        x = F.relu(self._linear(self.fc1, inputs))
        h_in = hidden_state.reshape(-1, self.args.hidden_dim)
        if self.args.use_rnn:
            h = self.rnn(x, h_in)
        else:
            h = F.relu(self._linear(self.rnn, x))
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
