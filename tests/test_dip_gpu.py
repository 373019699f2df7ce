"""Autograd contract of the drop-in boundary: `v` is a non-leaf tensor produced by a small convolutional
generator (the reference's deep-image-prior reparameterisation, ADFWI/dip/dip_acoustic_model.py:161); the gradient of a
record-space loss must reach the generator's weights through AcousticFD.backward + the torch coefficient algebra.
Golden: the unmodified reference kernel on CPU (tests/golden/make_golden_dip.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_gradients_reach_generator_weights(golden_dir):
    from adfwi_b200.propagator import acoustic_kernels as ak
    g = np.load(f"{golden_dir}/dip_acoustic.npz")
    dev = torch.device("cuda:0")
    net = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(8, 1, 3, padding=1))
    net.load_state_dict({k[2:]: torch.tensor(g[k]) for k in g.files if k.startswith("w_")})
    net = net.to(dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    t = lambda k: torch.tensor(g[k], device=dev)
    vp = 2500.0 + 800.0 * torch.tanh(net(t("z")))[0, 0]
    assert not vp.is_leaf
    assert rel_l2(vp.detach().cpu().numpy(), g["vp"]) < 1e-6
    rec = ak.forward_kernel(int(g["nx"]), int(g["nz"]), float(g["dx"]), float(g["dz"]), int(g["nt"]), float(g["dt"]), int(g["nabc"]), True,
                            t("src_x"), t("src_z"), 3, t("src_v"), t("rcv_x"), t("rcv_z"), len(g["rcv_x"]), t("damp"), vp, t("rho"),
                            checkpoint_segments=1, device=dev, dtype=torch.float32)
    assert rel_l2(rec["p"].detach().cpu().numpy(), g["rec_p"]) < 1e-5
    (rec["p"] * t("W")).sum().backward()
    for k, p in net.named_parameters():
        e = rel_l2(p.grad.cpu().numpy(), g["g_" + k])
        assert e < 1e-4, (k, e)
