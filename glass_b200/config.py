"""impl/config.py equivalent: process-wide device selection.  Only CUDA devices are usable."""
import torch

device = None


def set_device(idx):
    global device
    if idx == "cpu" or idx == -1:
        raise RuntimeError("glass_b200 has no CPU path: pass a CUDA device index (the reference's "
                           "`--device -1` CPU run is what bench.py --impl reference times)")
    if not torch.cuda.is_available():
        raise RuntimeError("glass_b200 needs a CUDA device")
    device = torch.device(f"cuda:{idx}")
    torch.cuda.set_device(device)
    return device
