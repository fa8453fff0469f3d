"""Measured baselines: ``ref_emulation`` — the reference's execution model re-created without its
infrastructure; ``nccl_torch`` — the honest library build (torch.func + cuDNN/cuBLAS + NCCL collectives)."""
