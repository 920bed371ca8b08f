"""BASELINE configs[4]: on-device render -> fused verifier pre-processing -> ResNet-152 early-fusion verifier (random-init weights),
everything resident on one GPU; prints one JSON line with hypotheses/s end to end and per stage.

    python scripts/bench_c5.py [--hyp 640] [--steps 3] [--dtype bf16|fp32] [--batch 256]

The verifier is the reference's architecture unchanged (salve/models/early_fusion.py:14-83): torchvision resnet152 whose conv1
takes 12 channels (x1c, x2c, x1f, x2f) and whose fc has 2 classes; it is a consumer of this repo's output, not part of it.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch import nn
from torchvision import models

from oracle import synth
from salve_b200.renderer import BevRenderer


class EarlyFusionCEResnet152(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnet = models.resnet152(weights=None)
        self.conv1 = nn.Conv2d(12, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.fc = nn.Linear(2048, 2)

    def forward(self, x):
        r = self.resnet
        x = r.maxpool(r.relu(r.bn1(self.conv1(x))))
        x = r.layer4(r.layer3(r.layer2(r.layer1(x))))
        return self.fc(torch.flatten(r.avgpool(x), 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hyp", type=int, default=640)
    ap.add_argument("--panos", type=int, default=40)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    rgbs, depths, p1, p2, R, t = synth.synth_building(a.panos, a.hyp, 512, 1024, seed=0)
    r = BevRenderer(max_panos=a.panos, max_images=1480)
    d_rgb = torch.from_numpy(rgbs).to(dev); d_dep = torch.from_numpy(depths.view(np.int16)).to(dev)
    for k in range(a.panos):
        r.bind_pano(k, d_rgb[k].data_ptr(), d_dep[k].data_ptr())
    ib = 501 * 501 * 3
    d_posed = torch.empty(a.hyp * 2 * ib, dtype=torch.uint8, device=dev)
    d_unposed = torch.empty(a.panos * 2 * ib, dtype=torch.uint8, device=dev)
    x = torch.empty((a.hyp, 12, 224, 224), dtype=torch.float32, device=dev)
    torch.manual_seed(0)
    model = EarlyFusionCEResnet152().to(dev).eval()
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    if dt != torch.float32:
        model = model.to(dt)
    model = model.to(memory_format=torch.channels_last)
    sh = torch.cuda.current_stream(dev).cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step():
        ev[0].record()
        idx, nu = r.render_hypotheses_compact_device(p1, p2, R, t, d_posed, d_unposed, stream=sh)
        ev[1].record()
        r.verifier_preprocess(r.quadruplet_pointers_compact(d_posed, d_unposed, idx), x, stream=sh)
        ev[2].record()
        outs = []
        with torch.no_grad():
            for b0 in range(0, a.hyp, a.batch):
                xb = x[b0 : b0 + a.batch].to(dt).contiguous(memory_format=torch.channels_last)
                outs.append(model(xb).float())
        logits = torch.cat(outs)
        ev[3].record()
        return logits

    for _ in range(2):
        logits = step()
    torch.cuda.synchronize()
    tot = np.zeros(3)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        logits = step()
        torch.cuda.synchronize()
        tot += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    wall = time.perf_counter() - t0
    host_logits = logits.cpu().numpy()
    print(json.dumps({
        "config": "C5: render -> preprocess -> ResNet-152 early fusion (random init), 1 GPU, device resident",
        "hypotheses": a.hyp, "steps": a.steps, "verifier_dtype": a.dtype, "verifier_batch": a.batch,
        "hyp_per_s_end_to_end": a.hyp * a.steps / wall,
        "ms_per_step": {"render": tot[0] / a.steps, "preprocess": tot[1] / a.steps, "resnet152": tot[2] / a.steps},
        "logits_finite": bool(np.isfinite(host_logits).all()), "logits_shape": list(host_logits.shape),
    }))


if __name__ == "__main__":
    main()
