"""Drop-in for ``/root/reference/utils/pad_tensor.py`` (host glue, unchanged semantics)."""
import torch
import torch.nn.functional as functional


def pad_tensor(inp):
    """Zero-pad a list of tensors to their common (element-wise maximum) shape."""
    assert type(inp[0]) == torch.Tensor
    max_shape = list(inp[0].shape)
    for t in inp[1:]:
        for i in range(len(max_shape)):
            max_shape[i] = int(max(max_shape[i], t.shape[i]))
    padded_ts = []
    for t in inp:
        pad_pattern = []
        for i in reversed(range(len(max_shape))):
            pad_pattern += [0, max_shape[i] - t.shape[i]]
        padded_ts.append(functional.pad(t, tuple(pad_pattern), 'constant', 0))
    return padded_ts


def pad_tensor_varied(inp, dummy=-100):
    """Pads with ``dummy`` to the common shape plus one in every dimension (pad_tensor.py:33-58)."""
    assert type(inp[0]) == torch.Tensor
    max_shape = list(inp[0].shape)
    for t in inp[1:]:
        for i in range(len(max_shape)):
            max_shape[i] = int(max(max_shape[i], t.shape[i]))
    max_shape = [m + 1 for m in max_shape]
    padded_ts = []
    for t in inp:
        pad_pattern = []
        for i in reversed(range(len(max_shape))):
            pad_pattern += [0, max_shape[i] - t.shape[i]]
        padded_ts.append(functional.pad(t, tuple(pad_pattern), 'constant', dummy))
    return padded_ts
