"""CPU: the product-side VAE decoder (torch / library convs) against the oracle restatement, and its
diffusers-compatible parameter names."""
import torch

from oracle import vae_decoder


def test_decoder_matches_oracle_restatement_fp32():
    from flite_b200 import vae
    torch.manual_seed(0)
    m = vae.AutoencoderKL().eval()
    ref = vae_decoder.Decoder().eval()
    missing, unexpected = ref.load_state_dict(vae_decoder.map_diffusers_names(m.state_dict()), strict=True)
    assert not missing and not unexpected
    z = torch.randn(2, 16, 8, 8)
    with torch.no_grad():
        a, b = m.decode(z).sample, ref(z)
    assert a.shape == (2, 3, 64, 64)
    assert ((a - b).norm() / b.norm()).item() <= 1e-5
    m.enable_slicing()
    assert torch.allclose(m.decode(z).sample, a, atol=1e-5)
    assert m.config.scaling_factor == vae_decoder.SCALING_FACTOR and m.config.shift_factor == vae_decoder.SHIFT_FACTOR


def test_decoder_parameter_names_follow_diffusers():
    from flite_b200 import vae
    keys = set(vae.AutoencoderKL().state_dict())
    for k in ("decoder.conv_in.weight", "decoder.mid_block.resnets.0.norm1.weight",
              "decoder.mid_block.attentions.0.group_norm.weight", "decoder.mid_block.attentions.0.to_q.weight",
              "decoder.mid_block.attentions.0.to_out.0.bias", "decoder.up_blocks.0.resnets.2.conv2.weight",
              "decoder.up_blocks.2.resnets.0.conv_shortcut.weight", "decoder.up_blocks.0.upsamplers.0.conv.weight",
              "decoder.up_blocks.3.resnets.0.conv_shortcut.weight", "decoder.conv_norm_out.weight",
              "decoder.conv_out.bias"):
        assert k in keys, k
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in keys        # last block does not upsample
    assert "decoder.up_blocks.1.resnets.0.conv_shortcut.weight" not in keys  # 512 -> 512
    n = sum(v.numel() for v in vae.AutoencoderKL().state_dict().values())
    assert 49_000_000 < n < 50_500_000                                       # FLUX VAE decoder: ~49.5 M parameters
