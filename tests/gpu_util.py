import numpy as np
import torch


def dev_graph(g, device="cuda", rowptr64=False):
    rp = torch.from_numpy(g.rowptr.astype(np.int64 if rowptr64 else g.rowptr.dtype)).to(device)
    return (torch.from_numpy(g.x).to(device), rp, torch.from_numpy(g.col).to(device), torch.from_numpy(g.ew).to(device))


def cpu_graph(g):
    return (torch.from_numpy(g.x), torch.from_numpy(g.rowptr), torch.from_numpy(g.col), torch.from_numpy(g.ew))


def assert_close_fp32(got, want, rtol=1e-4, atol_rms=1e-5, what=""):
    """The north star's tolerance: embeddings/rewards within 1e-4 relative (fp32).  Elementwise
    |got-want| <= rtol*|want| + atol_rms*rms(want): the absolute term (1e-5 of the tensor's RMS)
    covers entries that are themselves the result of cancellation."""
    got = got.detach().double().cpu(); want = want.detach().double().cpu()
    rms = float(want.pow(2).mean().sqrt()) if want.numel() else 0.0
    err = (got - want).abs()
    tol = rtol * want.abs() + atol_rms * max(rms, 1e-30)
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} out of tolerance, max err {float(err.max()):.3e} (rms {rms:.3e})"
