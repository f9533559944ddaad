"""Shape arithmetic in plain Python.  torch.broadcast_shapes goes through the meta / symbolic-shape machinery and costs
~30 us per call; the host path calls it a dozen times per objective step."""
import torch


def broadcast_shapes(*shapes):
    """Same result and same error type (RuntimeError) as torch.broadcast_shapes, as a torch.Size."""
    nd = 0
    for s in shapes:
        if len(s) > nd:
            nd = len(s)
    out = [1] * nd
    for s in shapes:
        off = nd - len(s)
        for i, d in enumerate(s):
            d = int(d)
            cur = out[off + i]
            if d == cur or d == 1:
                continue
            if cur == 1:
                out[off + i] = d
            else:
                raise RuntimeError("Shape mismatch: objects cannot be broadcast to a single shape. Mismatch is between "
                                   "arg with shape %s and the result so far %s" % (tuple(s), tuple(out)))
    return torch.Size(out)
