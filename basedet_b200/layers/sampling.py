"""Drop-in for ``basedet.layers.common.sampling`` (layers/common/sampling.py:7-30)."""
import torch

from .. import ops


def sample_labels(labels, num_samples, label_value, ignore_label=-1, noise=None):
    """Keep at most ``num_samples`` elements equal to ``label_value``; the surplus becomes ``ignore_label``.

    Same call as the reference (1-D ``labels``, modified in place and returned; bool masks as layers/head/rcnn.py:126
    passes them are returned as a new bool tensor).  The reference draws ``uniform(size=num_valid)`` from MegEngine's
    generator; here one uniform variate per element comes from ``noise`` (explicit, reproducible) or, if omitted, from
    torch's CUDA generator -- the selection rule on those variates is the reference's (largest variates go, ties by
    index)."""
    assert labels.ndim == 1, "Only tensor of dim 1 is supported."
    if noise is None:
        noise = torch.rand(labels.shape, dtype=torch.float32, device=labels.device)
    if labels.dtype == torch.bool:
        work = labels.to(torch.int32)
        ops.sample_labels(work, noise, int(num_samples), int(bool(label_value)), int(bool(ignore_label)))
        return work.to(torch.bool)
    if labels.dtype == torch.int32 and labels.is_contiguous():
        ops.sample_labels(labels, noise, int(num_samples), int(label_value), int(ignore_label))
        return labels
    work = labels.to(torch.int32).contiguous()
    ops.sample_labels(work, noise, int(num_samples), int(label_value), int(ignore_label))
    labels.copy_(work.to(labels.dtype))
    return labels
