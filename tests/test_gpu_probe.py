"""Hardware probe: UMMA SWIZZLE_128B descriptors with a start address shifted by whole rows.
Records the outcome in gpurun_out/umma_probe.json (it steers the conv kernel's halo-reuse design)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_umma_row_shift_probe():
    from tedm_b200 import native as N
    g = torch.Generator().manual_seed(0)
    A = torch.randn(384, 64, generator=g).to(torch.bfloat16).cuda()
    Bm = torch.randn(64, 64, generator=g).to(torch.bfloat16).cuda()
    shifts = [0, 8, 1, 1, 2, 2, 3, 3, 7, 7, 9, 9, 130, 130, 16, 66, 66]
    bos = [0, 0, 0, 1, 0, 2, 0, 3, 0, 7, 0, 1, 0, 2, 0, 0, 2]
    out = N.umma_probe(A, Bm, shifts, bos)
    torch.cuda.synchronize()
    res = []
    for v, (s, bo) in enumerate(zip(shifts, bos)):
        ref = A[s:s + 128].float() @ Bm.float().t()
        err = ((out[v] - ref).norm() / ref.norm()).item()
        res.append({"shift": s, "base_offset": bo, "rel_err": err, "ok": err < 1e-3})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "umma_probe.json"), "w"), indent=1)
    print(json.dumps(res))
    assert res[0]["ok"] and res[1]["ok"], "baseline UMMA (aligned start) is wrong"
