"""Drop-in for ``/root/reference/utils/feature_align.py`` (same three functions and arguments).

``feature_align`` is one kernel launch (``csrc/feature_align.cu``) instead of a python loop of ~40 tiny
ops per keypoint (``utils/feature_align.py:32-36,60-62,67-125``); results are bit-identical, including
the reference's (W,H) / (H_f,W_f) scaling mix-up and its post-fetch edge rule.
"""
import torch
from torch import Tensor

from fpmatch import ops


def feature_align(raw_feature: Tensor, P: Tensor, ns_t: Tensor, ori_size: tuple, device=None) -> Tensor:
    r"""
    :param raw_feature: :math:`(b\times c \times w \times h)` raw feature map
    :param P: :math:`(b\times n \times 2)` point set, coordinates at the scale of the original image
    :param ns_t: :math:`(b)` number of exact points
    :param ori_size: size of the original image
    :param device: output device. If not specified, it will be the same as the input
    :return: :math:`(b\times c \times n)` extracted feature vectors
    """
    if device is None:
        device = raw_feature.device
    ori = tuple(float(v) for v in (ori_size.tolist() if isinstance(ori_size, Tensor) else ori_size))
    F = ops.feature_align(raw_feature.detach().to(torch.float32).contiguous(),
                          P.to(raw_feature.device, torch.float32).contiguous(),
                          ns_t.to(raw_feature.device), ori)
    return F.to(device)


def interp_2d(z: Tensor, P: Tensor, ori_size: Tensor, feat_size: Tensor, out=None, device=None) -> Tensor:
    r"""
    Interpolate one feature map :math:`(c\times w\times h)` at the points ``P`` :math:`(n\times 2)`.
    ``feat_size`` is accepted for signature compatibility; like the reference's caller it must equal
    ``z.shape[1:3]`` (the kernel derives it from ``z``).
    """
    if device is None:
        device = z.device
    n = P.shape[0]
    ns = torch.tensor([n], dtype=torch.int64, device=z.device)
    res = feature_align(z.unsqueeze(0), P.unsqueeze(0), ns, ori_size)[0].to(device)
    if out is not None:
        out.copy_(res)
        return out
    return res


def bilinear_interpolate(im: Tensor, x: Tensor, y: Tensor, device=None):
    r"""
    Bi-linear interpolate a 3d feature map :math:`(c\times w\times h)` at feature-space coordinate (x, y).
    Same kernel as feature_align with the coordinate transform switched off.
    """
    if device is None:
        device = im.device
    xs = torch.as_tensor(x, dtype=torch.float32).reshape(())
    ys = torch.as_tensor(y, dtype=torch.float32).reshape(())
    P = torch.stack([xs, ys]).reshape(1, 1, 2).to(im.device)
    ns = torch.tensor([1], dtype=torch.int64, device=im.device)
    out = ops.feature_align(im.detach().to(torch.float32).contiguous().unsqueeze(0), P.contiguous(), ns,
                            (1.0, 1.0), feat_coords=True)
    return out[0, :, 0].to(device)
