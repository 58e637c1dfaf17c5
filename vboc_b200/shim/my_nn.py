"""Mirror of the reference's `my_nn.py:4-34`: the two MLP shapes every driver trains and the comparison scripts
re-create before `load_state_dict` (triplependulum_comparison.py:31-41).  Same class names, same
`linear_relu_stack` attribute (so state_dict keys are `linear_relu_stack.{0,2,4}.{weight,bias}`), same forward."""
import torch.nn as nn


def _stack(n_in, hidden, n_out, final_relu):
    layers = [nn.Linear(n_in, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(), nn.Linear(hidden, n_out)]
    return nn.Sequential(*(layers + ([nn.ReLU()] if final_relu else [])))


class NeuralNetCLS(nn.Module):
    """Classifier: logits out (the sigmoid is applied by the callers, AL/triplependulum_al.py:253-264)."""

    def __init__(self, input_size, hidden_size, output_size):
        super().__init__()
        self.linear_relu_stack = _stack(input_size, hidden_size, output_size, final_relu=False)

    def forward(self, x):
        return self.linear_relu_stack(x)


class NeuralNetDIR(nn.Module):
    """Regressor of the maximum velocity norm along a direction: non-negative output (final ReLU)."""

    def __init__(self, input_size, hidden_size, output_size):
        super().__init__()
        self.linear_relu_stack = _stack(input_size, hidden_size, output_size, final_relu=True)

    def forward(self, x):
        return self.linear_relu_stack(x)
