"""Profiling target: the VAE mid attention block (single head, d=512) on one frame of a 960x960 VAE tile (N = 120*120 = 14400
tokens): fused path (q|k|v GEMM + split-D flash kernel, csrc/attention_hd512.cu) by default, the panelled GEMM ->
softmax -> GEMM path with MGLD_VAE_FUSED_ATTN=0 (autoencoder._AttnBlock).  Run under
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --profile-from-start off --csv ...
Algorithmic traffic of the attention proper: Q + K + V + O = 4 * 14400 * 512 * 2 B = 59 MB (the N x N scores in fp32 + P in
fp16 would be 829 + 415 MB written and read again)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgld_vsr_b200 import ops
from mgld_vsr_b200.autoencoder import _AttnBlock
from mgld_vsr_b200.unet import _Packed
H = W = int(os.environ.get("MGLD_VAE_HW", "120"))
C = 512
g = torch.Generator().manual_seed(0)
sd = {}
for n in ("q", "k", "v", "proj_out"):
    sd[f"a.{n}.weight"] = torch.randn(C, C, 1, 1, generator=g) * C ** -0.5
    sd[f"a.{n}.bias"] = torch.zeros(C)
sd["a.norm.weight"], sd["a.norm.bias"] = torch.ones(C), torch.zeros(C)
blk = _AttnBlock(_Packed(sd, torch.device("cuda")), "a")
x = torch.randn(1, H, W, C, generator=g).half().cuda()
for k in range(2):
    if k == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    ops.stats_pool_reset()
    y = blk(ops, x)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done", y.float().abs().max().item())
