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
    Pd = P.to(raw_feature.device, torch.float32).contiguous()
    nsd = ns_t.to(raw_feature.device)
    if torch.is_grad_enabled() and raw_feature.requires_grad:
        # the reference backpropagates into the feature map through its slice assignments (feature_align.py:62)
        F = _FeatureAlignFn.apply(raw_feature.to(torch.float32), Pd, nsd, ori)
    else:
        F = ops.feature_align(raw_feature.detach().to(torch.float32).contiguous(), Pd, nsd, ori)
    return F.to(device)


class _FeatureAlignFn(torch.autograd.Function):
    """``feature_align`` with a gradient for the feature map (none for the keypoints, as in the reference, whose
    coordinates pass through ``floor`` / integer indexing).  Forward = the bit-exact kernel; backward = the transposed
    4-tap gather: every output column scatters its gradient to its four taps with the interpolation weights
    (``feature_align.py:98-125``), as one batched ``index_add`` on the device."""

    @staticmethod
    def forward(ctx, raw_feature, P, ns, ori):
        raw = raw_feature.contiguous()
        ctx.save_for_backward(P, ns)
        ctx.meta = (tuple(raw.shape), ori)
        return ops.feature_align(raw, P, ns, ori)

    @staticmethod
    def backward(ctx, g):
        P, ns = ctx.saved_tensors
        (B, C, Hf, Wf), ori = ctx.meta
        dev = g.device
        n = P.shape[1]
        ori_t = torch.tensor(ori, dtype=torch.float32, device=dev)
        feat = torch.tensor([Hf, Wf], dtype=torch.float32, device=dev)      # (sic) feature_align.py:30,57-61
        step = ori_t / feat
        pt = (P - step / 2) / ori_t * feat
        x, y = pt[..., 0], pt[..., 1]
        x0 = torch.floor(x); x1 = x0 + 1; y0 = torch.floor(y); y1 = y0 + 1
        x0 = x0.clamp(0, Wf - 1); x1 = x1.clamp(0, Wf - 1); y0 = y0.clamp(0, Hf - 1); y1 = y1.clamp(0, Hf - 1)
        xi0, xi1, yi0, yi1 = x0.long(), x1.long(), y0.long(), y1.long()
        eqx, eqy = xi0 == xi1, yi0 == yi1                                    # edge rule applied AFTER the fetch (:104-113)
        x0 = torch.where(eqx & (xi0 == 0), x0 - 1, x0); x1 = torch.where(eqx & (xi0 != 0), x1 + 1, x1)
        y0 = torch.where(eqy & (yi0 == 0), y0 - 1, y0); y1 = torch.where(eqy & (yi0 != 0), y1 + 1, y1)
        valid = (torch.arange(n, device=dev)[None, :] < ns.view(-1, 1)).to(g.dtype)
        w = [(x1 - x) * (y1 - y), (x1 - x) * (y - y0), (x - x0) * (y1 - y), (x - x0) * (y - y0)]
        taps = [yi0 * Wf + xi0, yi1 * Wf + xi0, yi0 * Wf + xi1, yi1 * Wf + xi1]
        d = torch.zeros((B, C, Hf * Wf), dtype=g.dtype, device=dev)
        for wk, tk in zip(w, taps):
            d.scatter_add_(2, tk[:, None, :].expand(B, C, n), g * (wk * valid)[:, None, :])
        return d.view(B, C, Hf, Wf), None, None, None


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
