"""CPU, build container only: oracle vs the LIVE unmodified reference on fresh random inputs
(skipped where /root/reference is not mounted, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import cbind, ref_loader, step_oracle as so

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ns():
    return ref_loader.load()


@pytest.mark.parametrize("seed", range(12))
def test_lrh_random(ns, seed):
    g = torch.Generator().manual_seed(seed)
    C = [6, 7, 3, 16][seed % 4]
    pct = [0.5, 0.9, 0.0, 0.3, 0.75][seed % 5]
    ign = [-1, 255][seed % 2]
    lab = torch.randint(0, C + 1, (3, 40, 24), generator=g)
    lab[torch.rand(3, 40, 24, generator=g) < 0.15] = ign
    reg = torch.randint(0, 9, (3, 40, 24), generator=g) * (seed % 3 + 1)
    ref = ns.Homogenizer(pct, C, ign)(lab, reg).numpy()
    assert np.array_equal(ref, cbind.lrh(lab.numpy(), reg.numpy(), C, ign, pct))


@pytest.mark.parametrize("seed", range(6))
def test_select_downscale_random(ns, seed):
    g = torch.Generator().manual_seed(100 + seed)
    soft = torch.softmax((seed + 1) * torch.randn(2, 6, 20, 28, generator=g), 1)
    for top, low in ((0.8, 0.6), (0.8, 0.2), (0.3, 0.1)):
        ref = ns.pseudo_selection(soft.clone(), top, low, "tensor", -1).numpy()
        assert np.array_equal(ref, cbind.pseudo_select(soft.numpy(), top, low, -1))
    lab = torch.randint(-1, 6, (2, 64, 48), generator=g)
    lab[:, : 16 * (seed % 4), :32] = seed % 6
    for scale, mr in ((16, 0.75), (8, 0.4)):
        ref = ns.DownscaleLabel(scale, 6, -1, mr)(lab).numpy()
        assert np.array_equal(ref, cbind.downscale_label(lab.numpy(), scale, 6, -1, mr))


def test_model_keys_and_forward(ns):
    ref = ref_loader.build_reference_model(ns, "resnet101", 6)
    mine = so.DeeplabOracle("resnet101", 6)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    assert len(ref.state_dict()) == 688                      # checkpoint ABI
    sd = so.seeded_state_dict(ref, 5)
    ref.load_state_dict(sd)
    mine.load_state_dict(sd)
    ref.eval(), mine.eval()
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        torch.testing.assert_close(ref(x), mine(x), rtol=1e-5, atol=1e-6)
