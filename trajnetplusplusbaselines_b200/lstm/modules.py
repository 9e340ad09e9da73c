"""Parameter containers with the reference's module names and state_dict keys.

Mirrors trajnetbaselines/lstm/modules.py (InputEmbedding :4-48, Hidden2Normal :51-64).  The
arithmetic of both is fused into lstm_gates_kernel (csrc/lstm_step.cu); these modules only own
the parameters so that reference checkpoints load verbatim.
"""
import torch

_FUSED = ("%s.forward is fused into the CUDA step kernel (csrc/lstm_step.cu); call LSTM.forward / "
          "LSTM.step -- there is no stand-alone torch path")


class InputEmbedding(torch.nn.Module):
    """Linear(input_dim, embedding_dim - 2) + ReLU + two zero tag channels, input scaled by `scale`."""

    def __init__(self, input_dim, embedding_dim, scale, use_tags=True):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.scale = scale
        self.use_tags = use_tags
        linear_embedding_dim = self.embedding_dim - 2 if use_tags else self.embedding_dim
        self.input_embeddings = torch.nn.Sequential(
            torch.nn.Linear(input_dim, linear_embedding_dim),
            torch.nn.ReLU(),
        )

    def forward(self, vel):
        raise NotImplementedError(_FUSED % "InputEmbedding")

    def _tag(self, vel, channel):
        # the two tag channels at the end of the embedding are reserved for start markers
        # (modules.py:32-48); the recurrence itself never uses them (they stay zero, A5)
        if not self.use_tags:
            raise Exception('Input embedding does not support start tag')
        tag = vel.new_zeros((vel.size(0), self.embedding_dim))
        tag[:, channel] = 1
        return tag

    def start_enc(self, vel):
        return self._tag(vel, -2)

    def start_dec(self, vel):
        return self._tag(vel, -1)


class Hidden2Normal(torch.nn.Module):
    """Linear(hidden_dim, 5) + sigmoid squashing of sigma_x, sigma_y, rho."""

    def __init__(self, hidden_dim):
        super().__init__()
        self.linear = torch.nn.Linear(hidden_dim, 5)

    def forward(self, hidden_state):
        raise NotImplementedError(_FUSED % "Hidden2Normal")
