"""Config 4 secondary numbers: the hot-path kernels of the BNN SGVB step (bnn_vi.py, layers [90,50,1], K=100 weight
particles) on synthetic UCI-shaped data.  GPU box only.
  * y-likelihood: Normal log-density of y[b] under mean[K,b], scalar std, summed over the batch -> [K]
    (8 B per particle-datapoint: read mean, write dmean), forward and backward;
  * weight nodes: sample + log q + log p of the 4601 weights per particle (fused latent kernels) and their backward.
The per-particle layers themselves are cuBLAS batched GEMMs (zhusuan.particle_linear) and are not claimed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
import torch
from zhusuan import _backend as be

FULL, KBCAST, SCALAR = be.FULL, be.KBCAST, be.SCALAR


def timeit(fn, reps=100):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


out = []
K = 100
for b in (8192, 131072):
    mean = torch.randn(K, b, device="cuda")
    y = torch.randn(b, device="cuda")
    std = torch.full((1,), 0.3, device="cuda")
    g = torch.randn(K, 1, device="cuda")
    # [K, M=1, E=b]: the batch is the event axis of the node (reduce over datapoints)
    tf = timeit(lambda: be.normal_logprob_fwd(y, KBCAST, mean, FULL, std, SCALAR, K, 1, b))
    tb = timeit(lambda: be.normal_logprob_bwd(g, y, KBCAST, mean, FULL, std, SCALAR, K, 1, b, False, True, False))
    n = K * b
    out.append(dict(kernel="y-likelihood forward (4 B per particle-datapoint)", K=K, batch=b, us=round(tf * 1e6, 2),
                    gbs=round(4 * n / tf / 1e9, 1), particle_datapoints_per_s=round(n / tf)))
    out.append(dict(kernel="y-likelihood backward (8 B per particle-datapoint)", K=K, batch=b, us=round(tb * 1e6, 2),
                    gbs=round(8 * n / tb / 1e9, 1), particle_datapoints_per_s=round(n / tb)))
    print(out[-2]); print(out[-1])
# weight nodes [K, 1, 4601] (50*91 + 1*51 weights per particle): sample, log q, log p through the general kernels
# (4601 is not a float4 multiple, so the public API does not take the fused latent form for them)
P = 4601
m, s_ = torch.zeros(P, device="cuda"), torch.ones(P, device="cuda")
ts = timeit(lambda: be.normal_sample(m, KBCAST, s_, KBCAST, K, P, seed=1, offset=4))
z = be.normal_sample(m, KBCAST, s_, KBCAST, K, P, seed=1, offset=4).reshape(K, 1, P)
tl = timeit(lambda: be.normal_logprob_fwd(z, FULL, m.reshape(1, P), KBCAST, s_.reshape(1, P), KBCAST, K, 1, P))
gk = torch.randn(K, 1, device="cuda")
tb = timeit(lambda: be.normal_logprob_bwd(gk, z, FULL, m.reshape(1, P), KBCAST, s_.reshape(1, P), KBCAST, K, 1, P, True,
                                          True, True))
out.append(dict(kernel="weight nodes: sample", K=K, weights=P, us=round(ts * 1e6, 2)))
out.append(dict(kernel="weight nodes: log-density forward", K=K, weights=P, us=round(tl * 1e6, 2)))
out.append(dict(kernel="weight nodes: log-density backward (x, mean, std)", K=K, weights=P, us=round(tb * 1e6, 2)))
for o in out[-3:]:
    print(o)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bnn_r1.json"), "w"), indent=1)
