"""Does torch.randn(..., out=view) consume the CUDA generator exactly like torch.randn(...)?  (dev check)"""
import torch
dev = "cuda"
torch.manual_seed(7)
a = [torch.randn(1024, 3, 32, 32, device=dev) for _ in range(3)]
torch.manual_seed(7)
buf = torch.empty(3, 1024, 3072, device=dev)
for i in range(3):
    torch.randn(1024, 3, 32, 32, device=dev, out=buf[i].view(1024, 3, 32, 32))
print("identical:", all(torch.equal(a[i].reshape(1024, -1), buf[i]) for i in range(3)))
