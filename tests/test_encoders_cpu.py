"""CPU: conditioning-encoder oracle against the reference goldens; drop-in surface (state_dict keys / shapes)."""
import numpy as np
import pytest
import torch

from helpers import DPP_CASES, ENCODER_CASES, build_encoder, dpp_case, encoder_case


@pytest.mark.parametrize("name", ENCODER_CASES)
def test_encoder_oracle_matches_reference_fp64(name):
    from oracle import encoders_oracle
    cls, kwargs, x, ref64, _, _ = encoder_case(name)
    enc = build_encoder(cls, kwargs)
    P = {k: v.double() for k, v in enc.state_dict().items()}
    if cls == "PhonemeEncoder":
        out = encoders_oracle.phoneme_encoder(P, x, heads=kwargs.get("heads", 8))
    else:
        out = encoders_oracle.speech_prompt_encoder(P, x.double(), heads=kwargs.get("heads", 8))
    assert np.abs(out.numpy() - ref64).max() < 1e-9


@pytest.mark.parametrize("name", ENCODER_CASES)
def test_encoder_state_dict_matches_reference(name):
    """Same keys, order and shapes as the reference module (recorded by make_golden.py from the reference itself)."""
    cls, kwargs, _, _, _, keys = encoder_case(name)
    enc = build_encoder(cls, kwargs)
    assert [(k, tuple(v.shape)) for k, v in enc.state_dict().items()] == [(k, tuple(s)) for k, s in keys]


def test_encoders_reject_cpu_inputs_and_masks():
    cls, kwargs, x, *_ = encoder_case("phon_small")
    enc = build_encoder(cls, kwargs)
    with pytest.raises(ValueError):
        enc(x)
    with pytest.raises(NotImplementedError):
        enc(x, mask=torch.ones_like(x, dtype=torch.bool))
    cls, kwargs, x, *_ = encoder_case("spe_small")
    with pytest.raises(ValueError):
        build_encoder(cls, kwargs)(x)


@pytest.mark.parametrize("name", DPP_CASES)
def test_duration_pitch_oracle_and_state_dict(name):
    from oracle import encoders_oracle
    kwargs, x, prompts, ref64, _, keys = dpp_case(name)
    enc = build_encoder("DurationPitchPredictor", kwargs)
    assert [(k, tuple(v.shape)) for k, v in enc.state_dict().items()] == [(k, tuple(s)) for k, s in keys]
    P = {k: v.double() for k, v in enc.state_dict().items()}
    dur, pitch = encoders_oracle.duration_pitch_predictor(P, x.double(), prompts.double(), heads=kwargs.get("heads", 8))
    assert np.abs(torch.stack((dur, pitch)).numpy() - ref64).max() < 1e-9


def test_length_regulation_oracle_and_host_glue_match_reference():
    """generate_mask_from_repeats / f0_to_coarse / expand_encodings: oracle == reference golden (bit-exact), and the
    product's torch glue (frame -> text index, coarse pitch bins) reproduces the reference's hard alignment."""
    from helpers import GOLDEN
    from naturalspeech2_pytorch_b200.encoders import f0_to_coarse, frames_to_text_index
    from oracle import encoders_oracle
    z = np.load(GOLDEN / "encoders.npz")
    ph, dur, pitch, table = (torch.from_numpy(z[f"expand_{k}"]) for k in ("phon", "duration", "pitch", "table"))
    np.testing.assert_array_equal(encoders_oracle.expand_encodings(ph, dur, pitch, table).numpy(), z["expand_cond"])
    idx = frames_to_text_index(dur)
    mask = encoders_oracle.generate_mask_from_repeats(dur)
    assert torch.equal(mask, idx.unsqueeze(1) == torch.arange(dur.shape[1]).view(1, -1, 1))
    assert torch.equal(f0_to_coarse(pitch), encoders_oracle.f0_to_coarse(pitch))
