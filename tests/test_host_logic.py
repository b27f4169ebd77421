"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the arena layout matches the reference's
state_dict contract, the drop-in surface has the reference's signatures and refuses to run without CUDA."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

from oracle import vae_oracle as vo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from vla_b200 import _lib
    header = open(os.path.join(ROOT, "include", "vla_b200.h")).read()
    declared = set(re.findall(r"\b(vla_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared
    assert lib.vla_abi_version() == 1


@pytest.mark.parametrize("kind", sorted(vo.MODEL_KINDS))
@pytest.mark.parametrize("dims", [dict(A=782, B=572, S=24, L=20, E=32), dict(A=1177, B=1211, S=24, L=37, E=64),
                                  dict(A=50, B=36, S=5, L=10, E=16)], ids=["baseline", "config_defaults", "small"])
def test_arena_layout_matches_state_dict_contract(kind, dims):
    from vla_b200 import Layout, _lib
    lay = Layout(kind, dims["A"], dims["B"], dims["S"], dims["L"], dims["E"])
    shapes = vo.param_shapes(kind, dims)
    got = {n: s for n, _, _, s in lay.entries}
    assert set(got) == set(shapes)
    for n, s in shapes.items():
        assert tuple(got[n]) == tuple(s), n
    # parameters do not overlap, are 16-byte aligned, and fit the arena
    spans = sorted((off, off + int(np.prod(shape))) for n, k, off, shape in lay.entries if k == _lib.TENSOR_PARAM)
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    assert spans[-1][1] <= lay.n_params
    # fused groups: fc_mu / fc_logvar rows are adjacent (one [2L, in] GEMM); decoder first layers are adjacent
    for prefix, _ in ([] if vo.is_ae(kind) else vo.MODEL_KINDS[kind]["encoders"]):     # (autoencoders: one head per encoder)
        mu_off, mu_shape = lay.params[prefix + ".fc_mu.weight"]
        lv_off, _ = lay.params[prefix + ".fc_logvar.weight"]
        assert lv_off == mu_off + mu_shape[0] * mu_shape[1]
    decs = [p for p, _ in vo.MODEL_KINDS[kind]["decoders"]]
    for a, b in zip(decs, decs[1:]):
        off_a, shape_a = lay.params[a + ".fc.0.weight"]
        assert lay.params[b + ".fc.0.weight"][0] == off_a + shape_a[0] * shape_a[1]
    # pack / unpack round trip
    named = {n: torch.randn(*s) for n, (o, s) in lay.params.items()}
    flat = lay.pack(named, extra=4)
    assert flat.numel() == lay.n_params + 4
    back = lay.unpack(flat)
    for n in named:
        assert torch.equal(back[n], named[n]), n


def test_dropin_surface_signatures_and_init():
    from src.models import (DNA2RNAVAE, DecoderA, DecoderB, DecoderC, EncoderA, EncoderB, EncoderC, MultiModalVAE, RNA2DNAVAE,
                            reparameterize)
    from src.utils import vae_loss
    from src.utils.directional_losses import dna2rna_loss, rna2dna_loss
    assert list(inspect.signature(MultiModalVAE.__init__).parameters)[1:] == ["input_dim_a", "input_dim_b", "n_sites", "latent_dim", "embed_dim"]
    assert list(inspect.signature(MultiModalVAE.forward).parameters)[1:] == ["a", "b", "site"]
    assert list(inspect.signature(RNA2DNAVAE.forward).parameters)[1:] == ["rna", "site"]
    assert list(inspect.signature(DNA2RNAVAE.forward).parameters)[1:] == ["dna", "site"]
    assert list(inspect.signature(vae_loss).parameters) == ["recon_a", "a", "recon_b", "b", "recon_c", "site", "mu", "logvar", "beta",
                                                            "gamma", "class_weights"]
    assert list(inspect.signature(rna2dna_loss).parameters) == ["recon_dna", "dna", "mu", "logvar", "beta"]
    assert list(inspect.signature(dna2rna_loss).parameters) == ["recon_rna", "rna", "mu", "logvar", "beta"]
    m = MultiModalVAE(782, 572, 24, 20)
    assert sum(p.numel() for p in m.parameters()) == 1081114
    assert isinstance(m.encoder_a, EncoderA) and isinstance(m.encoder_b, EncoderB) and isinstance(m.encoder_c, EncoderC)
    assert isinstance(m.decoder_a, DecoderA) and isinstance(m.decoder_b, DecoderB) and isinstance(m.decoder_c, DecoderC)
    w = m.encoder_a.fc[0].weight if hasattr(m.encoder_a.fc, "__getitem__") else dict(m.named_parameters())["encoder_a.fc.0.weight"]
    assert float(w.abs().max()) <= 1 / np.sqrt(782) + 1e-7                       # nn.Linear default init bound
    sd = m.state_dict()
    assert torch.all(sd["encoder_a.fc.1.running_var"] == 1) and int(sd["encoder_a.fc.1.num_batches_tracked"]) == 0
    assert m() == (None, None, None, None, None)
    assert RNA2DNAVAE(50, 36, 5, 8)() == (None, None, None)
    z = reparameterize(torch.zeros(4, 3), torch.zeros(4, 3))
    assert z.shape == (4, 3)
    # state_dict round trip through torch.save-compatible tensors
    m2 = MultiModalVAE(782, 572, 24, 20)
    m2.load_state_dict(sd)
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_autoencoder_dropin_surface():
    """RNA2DNAAE / DNA2RNAAE and their losses keep the reference's names, signatures, return arity and state_dict keys
    (src/models/directional_ae.py:17-59, 76-123; src/utils/ae_losses.py:8-39)."""
    from src.models.directional_ae import DNA2RNAAE, RNA2DNAAE
    from src.utils.ae_losses import dna2rna_ae_loss, rna2dna_ae_loss
    assert list(inspect.signature(RNA2DNAAE.__init__).parameters)[1:] == ["rna_dim", "dna_dim", "n_sites", "latent_dim", "embed_dim"]
    assert list(inspect.signature(RNA2DNAAE.forward).parameters)[1:] == ["rna", "site"]
    assert list(inspect.signature(DNA2RNAAE.forward).parameters)[1:] == ["dna", "site"]
    assert list(inspect.signature(rna2dna_ae_loss).parameters) == ["recon_dna", "dna"]
    assert list(inspect.signature(dna2rna_ae_loss).parameters) == ["recon_rna", "rna"]
    dims = dict(A=782, B=572, S=24, L=20, E=32)
    for cls, kind in ((RNA2DNAAE, "rna2dna_ae"), (DNA2RNAAE, "dna2rna_ae")):
        m = cls(782, 572, 24, 20)
        shapes = vo.param_shapes(kind, dims)
        sd = m.state_dict()
        assert list(sd) == list(shapes)                                       # same keys in the same order
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(shapes[k]), k
        assert m() == (None, None)
        with pytest.raises(RuntimeError, match="CPU"):
            m(site=torch.zeros(4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CPU"):
        rna2dna_ae_loss(torch.rand(4, 36), torch.rand(4, 36))


def test_population_shard_partitions_exactly():
    """vla_b200.shard: every member goes to exactly one rank (replicas only, no communication)."""
    from vla_b200 import shard
    items = list(range(23))
    for world in (1, 2, 4, 8):
        parts = [shard(items, r, world) for r in range(world)]
        assert sorted(x for p in parts for x in p) == items
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_no_cpu_fallback_anywhere():
    from src.models import DNA2RNAVAE
    from src.utils.directional_losses import dna2rna_loss
    m = DNA2RNAVAE(50, 36, 5, 8)
    with pytest.raises(RuntimeError, match="CPU"):
        m(dna=torch.rand(4, 36), site=torch.zeros(4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CPU"):
        dna2rna_loss(torch.rand(4, 50), torch.rand(4, 50), torch.zeros(4, 8), torch.zeros(4, 8))
    with pytest.raises(RuntimeError):
        m.encoder_dna(torch.rand(4, 36))       # stacks are parameter containers, not a slow path


def test_dataset_matches_reference_behaviour():
    import pandas as pd
    from torch.utils.data import DataLoader
    from src.data import MultiModalDataset
    rng = np.random.default_rng(0)
    tpm = rng.random((10, 7), dtype=np.float32)
    beta = rng.random((10, 5), dtype=np.float32)
    site = rng.integers(0, 3, 10)
    df = pd.DataFrame({"tpm_unstranded": list(tpm), "beta_value": list(beta), "primary_site_encoded": site})
    ds = MultiModalDataset(df)
    assert len(ds) == 10 and ds.tpm_data.dtype == np.float32 and ds.primary_site.dtype == np.int64
    t, b, s = ds[3]
    assert torch.equal(t, torch.tensor(tpm[3])) and torch.equal(b, torch.tensor(beta[3])) and int(s) == site[3]
    assert s.dtype == torch.long and s.dim() == 0
    ds2 = MultiModalDataset.from_numpy(tpm, beta, site)
    assert np.array_equal(ds2.beta_data, beta)
    batches = list(DataLoader(ds, batch_size=4, shuffle=False))
    assert [x[0].shape[0] for x in batches] == [4, 4, 2]
    assert torch.equal(batches[1][0], torch.tensor(tpm[4:8])) and torch.equal(batches[2][2], torch.tensor(site[8:10]))
    assert len(list(DataLoader(ds, batch_size=4, shuffle=True, drop_last=True))) == 2
