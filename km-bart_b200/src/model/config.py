"""MultiModalBartConfig — drop-in for the reference's src/model/config.py:4-92.

The reference subclasses transformers-3.0.2 BartConfig; this one is self-contained (no
transformers import) but keeps the same constructor arguments, defaults, attribute names,
JSON round trip (`from_dict` / `to_dict` / `from_pretrained` / `save_pretrained`) and the
PretrainedConfig generation defaults that GenerationMixin.generate reads
(src/model/mixins.py:150-173)."""
import copy
import json
import os

CONFIG_NAME = "config.json"


class MultiModalBartConfig:
    model_type = "bart"

    def __init__(
            self,
            activation_dropout=0.0,
            extra_pos_embeddings=2,
            activation_function="gelu",
            vocab_size=50320,
            image_feature_size=2048 + 4,
            d_model=1024,
            encoder_ffn_dim=4096,
            encoder_layers=12,
            encoder_attention_heads=16,
            decoder_ffn_dim=4096,
            decoder_layers=12,
            decoder_attention_heads=16,
            encoder_layerdrop=0.0,
            decoder_layerdrop=0.0,
            attention_dropout=0.0,
            dropout=0.1,
            max_position_embeddings=1024,
            init_std=0.02,
            classif_dropout=0.0,
            num_labels=1,
            num_attributes=1,
            num_relations=1,
            is_encoder_decoder=True,
            pad_token_id=1,
            bos_token_id=0,
            eos_token_id=2,
            img_feat_id=50273,
            cls_token_id=50276,
            normalize_before=False,
            add_final_layer_norm=False,
            scale_embedding=False,
            normalize_embedding=True,
            static_position_embeddings=False,
            add_bias_logits=False,
            decoder_start_token_id=0,
            partial_load=(),
            lm_loss_factor=1.0,
            mrm_loss_factor=1.0,
            attribute_loss_factor=1.0,
            relation_loss_factor=1.0,
            **common_kwargs
    ):
        # PretrainedConfig defaults (HF-3.0.2 configuration_utils) that the path reads
        self.output_attentions = common_kwargs.pop("output_attentions", False)
        self.output_hidden_states = common_kwargs.pop("output_hidden_states", False)
        self.use_cache = common_kwargs.pop("use_cache", True)
        self.max_length = common_kwargs.pop("max_length", 20)
        self.min_length = common_kwargs.pop("min_length", 0)
        self.do_sample = common_kwargs.pop("do_sample", False)
        self.early_stopping = common_kwargs.pop("early_stopping", False)
        self.num_beams = common_kwargs.pop("num_beams", 1)
        self.temperature = common_kwargs.pop("temperature", 1.0)
        self.top_k = common_kwargs.pop("top_k", 50)
        self.top_p = common_kwargs.pop("top_p", 1.0)
        self.repetition_penalty = common_kwargs.pop("repetition_penalty", 1.0)
        self.length_penalty = common_kwargs.pop("length_penalty", 1.0)
        self.no_repeat_ngram_size = common_kwargs.pop("no_repeat_ngram_size", 0)
        self.bad_words_ids = common_kwargs.pop("bad_words_ids", None)
        self.num_return_sequences = common_kwargs.pop("num_return_sequences", 1)
        self.model_type = common_kwargs.pop("model_type", "bart")

        self.activation_dropout = activation_dropout
        self.extra_pos_embeddings = extra_pos_embeddings
        self.activation_function = activation_function
        self.vocab_size = vocab_size
        self.d_model = d_model
        self.encoder_ffn_dim = encoder_ffn_dim
        self.encoder_layers = self.num_hidden_layers = encoder_layers
        self.encoder_attention_heads = encoder_attention_heads
        self.decoder_ffn_dim = decoder_ffn_dim
        self.decoder_layers = decoder_layers
        self.decoder_attention_heads = decoder_attention_heads
        self.encoder_layerdrop = encoder_layerdrop
        self.decoder_layerdrop = decoder_layerdrop
        self.attention_dropout = attention_dropout
        self.dropout = dropout
        self.max_position_embeddings = max_position_embeddings
        self.init_std = init_std
        self.classif_dropout = classif_dropout
        self.num_labels = num_labels
        self.is_encoder_decoder = is_encoder_decoder
        self.pad_token_id = pad_token_id
        self.bos_token_id = bos_token_id
        self.eos_token_id = eos_token_id
        self.normalize_before = normalize_before
        self.add_final_layer_norm = add_final_layer_norm
        self.scale_embedding = scale_embedding
        self.normalize_embedding = normalize_embedding
        self.static_position_embeddings = static_position_embeddings
        self.add_bias_logits = add_bias_logits
        self.decoder_start_token_id = decoder_start_token_id

        self.image_feature_size = image_feature_size
        self.img_feat_id = img_feat_id
        self.cls_token_id = cls_token_id
        self.partial_load = partial_load
        self.num_attributes = num_attributes
        self.num_relations = num_relations
        self.lm_loss_factor = lm_loss_factor
        self.mrm_loss_factor = mrm_loss_factor
        self.attribute_loss_factor = attribute_loss_factor
        self.relation_loss_factor = relation_loss_factor
        for k, v in common_kwargs.items():   # unknown keys are kept as attributes, like PretrainedConfig
            setattr(self, k, v)

    # BartConfig properties used by callers
    @property
    def num_attention_heads(self):
        return self.encoder_attention_heads

    @property
    def hidden_size(self):
        return self.d_model

    @classmethod
    def from_dict(cls, config_dict, **kwargs):
        d = dict(config_dict)
        d.pop("num_hidden_layers", None)
        d.update(kwargs)
        return cls(**d)

    def to_dict(self):
        out = copy.deepcopy(self.__dict__)
        out["model_type"] = self.model_type
        if isinstance(out.get("partial_load"), tuple):
            out["partial_load"] = list(out["partial_load"])
        return out

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    @classmethod
    def from_json_file(cls, path):
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_dict(json.load(f))

    @classmethod
    def from_pretrained(cls, path, **kwargs):
        if os.path.isdir(path):
            path = os.path.join(path, CONFIG_NAME)
        if not os.path.isfile(path):
            raise EnvironmentError(f"Can't load config from '{path}' (no network access: only local files/directories)")
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_dict(json.load(f), **kwargs)

    def save_pretrained(self, save_directory):
        assert os.path.isdir(save_directory), "Saving path should be a directory where the model and configuration can be saved"
        with open(os.path.join(save_directory, CONFIG_NAME), "w", encoding="utf-8") as f:
            f.write(self.to_json_string())

    def __repr__(self):
        return f"{self.__class__.__name__} {self.to_json_string()}"
