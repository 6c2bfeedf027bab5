import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sodt_b200.runtime import Detector
g = torch.Generator().manual_seed(0)
for (B, S) in ((5, 512), (1, 1536), (2, 1024)):
    rgb = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).cuda()
    ir = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).cuda()
    outs = []
    for dt, graph in ((torch.bfloat16, True), (torch.float32, False)):
        det = Detector(device="cuda", dtype=dt, seed=1, conf_thres=1e-6, cuda_graph=graph)
        p = det.predict(rgb, ir).float()
        buf = det.detect_device(rgb, ir)
        torch.cuda.synchronize()
        outs.append(p)
        del det
    rel = ((outs[0] - outs[1]).norm() / outs[1].norm()).item()
    print(f"B={B} S={S}: pred {tuple(outs[0].shape)} bf16 vs fp32 rel diff {rel:.3e}, counts {buf.counts.tolist()[:5]}")
    assert rel < 2e-2
print("ok")
