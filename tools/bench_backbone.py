"""How long does the stock ResNet-18 backbone (out of scope) take next to the matching head?
Times the two backbone chunks of Net on 2 x B images of 240 x 320 (fp32 module, cuDNN defaults) in NCHW, in
channels-last, and channels-last replayed from a CUDA graph.   python tools/bench_backbone.py [B]"""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "fingerprint-matching-code_b200"), str(ROOT)]
import torch
from src.model.ngm import Net

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
net = Net(regression=True).to("cuda").eval()
img = torch.randn(2 * B, 3, 240, 320, device="cuda")


def timed(fn, warm=3, steps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def fwd(x):
    with torch.no_grad():
        n = net.node_layers(x)
        return n, net.edge_layers(n)


res = {"images": 2 * B, "allow_tf32_cudnn": torch.backends.cudnn.allow_tf32}
res["nchw_ms"] = timed(lambda: fwd(img))
net_cl = net.to(memory_format=torch.channels_last)
img_cl = img.contiguous(memory_format=torch.channels_last)
res["channels_last_ms"] = timed(lambda: fwd(img_cl))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        fwd(img_cl)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = fwd(img_cl)
res["channels_last_graph_ms"] = timed(g.replay)
with torch.autocast("cuda", dtype=torch.bfloat16):
    res["channels_last_bf16_autocast_ms"] = timed(lambda: fwd(img_cl))
print(json.dumps(res))
