"""CPU suite: the drop-in boundary.  No kernel is launched here (no GPU in the build container):
  * libkmbart_sm100.so loads and exports every symbol include/kmbart.h declares;
  * the ctypes prototypes cover exactly those symbols;
  * the product `src.model` classes reproduce the reference's state_dict keys / shapes /
    parameters() order, config JSON round trip, checkpoint IO incl.
    partial_load, and error conventions;
  * there is no CPU fallback: compute entry points raise on a CPU model."""
import ctypes
import json
import os
import re
import subprocess

import pytest
import torch

from oracle import kmbart_oracle as O
import golden_cases as G
from helpers import product_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kmbart.h")


@pytest.fixture(scope="session")
def built_lib():
    from kmbart import lib as L
    if not os.path.exists(L.LIB_PATH):
        subprocess.run(["make", "-j8", "-C", ROOT, "km-bart_b200/libkmbart_sm100.so"], check=True)
    return L


def header_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kmb_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_documented_ops():
    syms = header_symbols()
    for s in ("kmb_gemm", "kmb_attn_fwd", "kmb_attn_bwd", "kmb_embed_ln_fwd", "kmb_layernorm_fwd", "kmb_layernorm_bwd",
              "kmb_ce_combine", "kmb_adamw_multi", "kmb_arch_check", "kmb_version", "kmb_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_prototypes_cover_the_header(built_lib):
    assert sorted(built_lib.EXPORTED_SYMBOLS) == header_symbols()
    lib = built_lib.load()
    assert lib.kmb_version() >= 100


def test_library_is_sm100a_native_code(built_lib):
    """tcgen05 / TMA must be in the shipped SASS (UTC*MMA, UTMALDG), and only sm_100a code."""
    out = subprocess.run(["cuobjdump", "-lelf", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_arch_check_fails_loudly_without_b200(built_lib):
    if torch.cuda.is_available() and torch.cuda.get_device_capability()[0] == 10:
        pytest.skip("running on a B200")
    with pytest.raises(built_lib.KmbartError):
        built_lib.check(built_lib.load().kmb_arch_check(), "kmb_arch_check")


# ------------------------------------------------------------------ model surface
@pytest.fixture(scope="module")
def small_model():
    from src.model.model import MultiModalBartForConditionalGeneration
    ocfg = G.small_config()
    return ocfg, MultiModalBartForConditionalGeneration(product_config(ocfg))


def test_state_dict_keys_shapes_and_param_order(small_model):
    ocfg, model = small_model
    sd = O.full_state_dict(O.init_state_dict(ocfg))
    got = model.state_dict()
    assert set(got) == set(sd)
    for k in sd:
        assert tuple(got[k].shape) == tuple(sd[k].shape), k
    assert [n for n, _ in model.named_parameters()] == list(O.param_shapes(ocfg))
    assert model.model.encoder.embed_tokens.weight is model.model.shared.weight is model.model.decoder.embed_tokens.weight


def test_pretraining_model_param_order_and_heads():
    from src.model.model import MultiModalBartForPreTraining
    ocfg = G.small_config(num_labels=1601, num_attributes=129, num_relations=129)
    model = MultiModalBartForPreTraining(product_config(ocfg))
    assert [n for n, _ in model.named_parameters()] == list(O.param_shapes(ocfg, pretraining=True))
    assert model.relation_head.dense.weight.shape == (128, 256)


def test_base_config_json_matches_reference_fields():
    from src.model.config import MultiModalBartConfig
    with open(os.path.join(ROOT, "configs", "vcg_base.json")) as f:
        d = json.load(f)
    cfg = MultiModalBartConfig.from_dict(d)
    assert (cfg.d_model, cfg.encoder_layers, cfg.decoder_layers, cfg.vocab_size) == (768, 6, 6, 50320)
    assert cfg.image_feature_size == 2052 and cfg.img_feat_id == 50273 and cfg.cls_token_id == 50276
    assert "model.shared.weight" in cfg.partial_load
    # PretrainedConfig generation defaults read by generate() (src/model/mixins.py:150-173)
    assert (cfg.max_length, cfg.num_beams, cfg.top_k, cfg.use_cache) == (20, 1, 50, True)
    rt = MultiModalBartConfig.from_dict(json.loads(cfg.to_json_string()))
    assert rt.to_dict() == cfg.to_dict()


def test_save_and_from_pretrained_round_trip(tmp_path, small_model):
    from src.model.model import MultiModalBartForConditionalGeneration
    ocfg, model = small_model
    model.save_pretrained(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["config.json", "pytorch_model.bin"]
    saved = torch.load(tmp_path / "pytorch_model.bin")
    assert "model.encoder.embed_tokens.weight" in saved and "final_logits_bias" in saved   # tied weights saved 3x
    m2 = MultiModalBartForConditionalGeneration.from_pretrained(str(tmp_path))
    for (n1, p1), (n2, p2) in zip(model.named_parameters(), m2.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)
    assert not m2.training   # from_pretrained returns eval() like the reference (src/model/mixins.py:864-867)


def test_partial_load_slice_copy(tmp_path, small_model):
    """src/model/mixins.py:511-528 — a vocab-50265 checkpoint loads into the vocab-50320 model."""
    from src.model.model import MultiModalBartForConditionalGeneration
    from src.model.config import MultiModalBartConfig
    ocfg, model = small_model
    ck = {k: v.clone() for k, v in model.state_dict().items()}
    for k in ("model.shared.weight", "model.encoder.embed_tokens.weight", "model.decoder.embed_tokens.weight"):
        ck[k] = ck[k][:50265] + 1.0
    ck["final_logits_bias"] = ck["final_logits_bias"][:, :50265] + 1.0
    torch.save(ck, tmp_path / "pytorch_model.bin")
    d = product_config(ocfg).to_dict()
    d["partial_load"] = ["final_logits_bias", "model.shared.weight", "model.encoder.embed_tokens.weight",
                         "model.decoder.embed_tokens.weight"]
    cfg = MultiModalBartConfig.from_dict(d)
    m2, info = MultiModalBartForConditionalGeneration.from_pretrained(str(tmp_path), config=cfg, error_on_mismatch=False,
                                                                      output_loading_info=True)
    w = m2.model.shared.weight
    assert torch.equal(w[:50265], ck["model.shared.weight"])
    assert torch.equal(m2.final_logits_bias[:, :50265], ck["final_logits_bias"]) and (m2.final_logits_bias[:, 50265:] == 0).all()
    assert not info["error_msgs"]
    # without partial_load the mismatch is logged, never raised (src/model/mixins.py:856-863)
    m3, info3 = MultiModalBartForConditionalGeneration.from_pretrained(str(tmp_path), config=product_config(ocfg),
                                                                       output_loading_info=True)
    assert info3["error_msgs"]


def test_from_pretrained_missing_path_raises(tmp_path):
    from src.model.model import MultiModalBartForConditionalGeneration
    with pytest.raises(EnvironmentError):
        MultiModalBartForConditionalGeneration.from_pretrained(str(tmp_path / "nope"), config=product_config(G.small_config()))


def test_no_cpu_fallback(small_model):
    ocfg, model = small_model
    batch = O.synthetic_batch(ocfg, batch=2, n_regions=3, n_ctx=8, tgt_len=5, seed=1)
    with pytest.raises(RuntimeError):
        model(**batch)
    with pytest.raises(RuntimeError):
        model.generate(input_ids=batch["input_ids"], image_features=batch["image_features"], max_length=4)
    with pytest.raises(NotImplementedError):
        model.model.decoder(batch["decoder_input_ids"])


def test_pretraining_forward_argument_errors():
    from src.model.model import MultiModalBartForPreTraining
    pcfg, _, pbatch = G.case_pretrain()
    model = MultiModalBartForPreTraining(product_config(pcfg))
    bad = dict(pbatch, mrm_mask=None)
    with pytest.raises(ValueError):
        model(**bad)


def test_generate_argument_asserts(small_model):
    ocfg, model = small_model
    ids = torch.zeros(2, 4, dtype=torch.long)
    with pytest.raises(AssertionError):
        model.generate(input_ids=ids, image_features=[], max_length=0)
    with pytest.raises(AssertionError):
        model.generate(input_ids=ids, image_features=[], num_beams=1, num_return_sequences=2)


def test_adamw_argument_validation():
    from kmbart.optim import AdamW
    p = [torch.nn.Parameter(torch.zeros(3))]
    with pytest.raises(ValueError):
        AdamW(p, lr=-1.0)
    with pytest.raises(ValueError):
        AdamW(p, betas=(1.0, 0.999))
    opt = AdamW(p)
    assert opt.defaults["eps"] == 1e-6 and opt.defaults["weight_decay"] == 0.0 and opt.defaults["correct_bias"] is True


def test_param_store_layout(small_model):
    """Flat master/grad buffers: q|k|v (and cross k|v) adjacent, 256-byte aligned slots, values preserved."""
    from kmbart.engine import ParamStore
    from src.model.model import MultiModalBartForConditionalGeneration
    ocfg = G.small_config()
    model = MultiModalBartForConditionalGeneration(product_config(ocfg))
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    st = ParamStore(model)
    d = ocfg.d_model
    for n, p in model.named_parameters():
        assert torch.equal(p, before[n])
        assert p.data_ptr() == st.P.data_ptr() + 4 * st.offsets[n]
    for lp in ("model.encoder.layers.0.self_attn", "model.decoder.layers.1.self_attn"):
        o = st.offsets[lp + ".q_proj.weight"]
        assert st.offsets[lp + ".k_proj.weight"] == o + d * d and st.offsets[lp + ".v_proj.weight"] == o + 2 * d * d
    o = st.offsets["model.decoder.layers.0.encoder_attn.k_proj.weight"]
    assert st.offsets["model.decoder.layers.0.encoder_attn.v_proj.weight"] == o + d * d
    assert st.offsets["model.shared.weight"] == st.zero_end and st.is_adopted()


def test_packed_image_features_is_list_like():
    """kmbart.feed.PackedImageFeatures stands in for the reference's list of per-sample RoI tensors
    (src/data/collation.py:68-213 output, src/training.py:120-130 consumer): len / index / iteration give views."""
    import torch
    from kmbart.feed import PackedImageFeatures, unwrap_features
    t = torch.arange(5 * 2052, dtype=torch.float32).view(5, 2052)
    pf = PackedImageFeatures(t, [2, 0, 3])
    assert len(pf) == 3
    parts = pf.as_list()
    assert [p.shape[0] for p in parts] == [2, 0, 3]
    assert parts[2].data_ptr() == t[2:].data_ptr() and torch.equal(pf[0], t[:2])
    ft, counts = unwrap_features(pf)
    assert ft is t and counts == [2, 0, 3]
    lst = [t[:2], t[2:]]
    assert unwrap_features(lst) == (lst, None)
    import pytest
    with pytest.raises(AssertionError):
        PackedImageFeatures(t, [1, 1])
