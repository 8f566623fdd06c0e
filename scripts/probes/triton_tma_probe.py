"""Debug probe (not part of the product): does a Triton TMA tensor-descriptor load work on this box, and what PTX / SASS does it emit?"""
import torch, triton, triton.language as tl
from triton.tools.tensor_descriptor import TensorDescriptor

@triton.jit
def k(desc, out_ptr, y, x, BH: tl.constexpr, BW: tl.constexpr):
    t = desc.load([y, x])
    offs = tl.arange(0, BH)[:, None] * BW + tl.arange(0, BW)[None, :]
    tl.store(out_ptr + offs, t)

img = (torch.arange(480 * 640, device="cuda", dtype=torch.int32) * 7 % 65000).to(torch.int16).reshape(480, 640)
desc = TensorDescriptor.from_tensor(img, [16, 64])
out = torch.zeros(16 * 64, device="cuda", dtype=torch.int16)
h = k[(1,)](desc, out, 50, 96, BH=16, BW=64)
torch.cuda.synchronize()
print("triton TMA load ok:", bool((out.reshape(16, 64) == img[50:66, 96:160]).all()))
ptx = h.asm["ptx"]
for line in ptx.splitlines():
    if "cp.async.bulk" in line or "mbarrier" in line or ".target" in line or ".version" in line:
        print(line.strip()[:200])
import re
sass = h.asm.get("sass", "")
for line in sass.splitlines():
    if "UTMALDG" in line or "UTMAPF" in line:
        print(line.strip()[:160])
