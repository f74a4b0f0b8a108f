"""FastMultimodalRecommender: host-side mirror of the reference
``MultimodalRecommender`` (reference ``src/models/multimodal.py:31-610``) for
the scoring path.  Same constructor arguments, same ``forward`` keyword
arguments, same ``state_dict`` key names and shapes, so a reference checkpoint
(``torch.load(path)['model_state_dict']``, ``scripts/evaluate.py:366-375``)
loads unchanged; the arithmetic runs in libpxr.so on the GPU.

Differences, all deliberate:
  * the frozen HF backbones are hoisted out: ``image`` is the cached (B, Dv)
    vision feature and ``text_input_ids`` the cached (B, Dl) text feature (what
    the backbones would emit); backbone / contrastive-head keys of a checkpoint
    are accepted and ignored;
  * ``fusion_type='attention'`` follows the documented semantics of
    ``AttentionFusionLayer`` (``layers.py:135-164``); the reference's own call
    raises a TypeError;
  * ``forward`` is inference-only (eval mode, no autograd).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .engine import PxrEngine, check_index_range

# dims of the reference's backbone registry (reference src/config.py:18-31); 'cached<N>' is the
# hoisted-backbone spelling used by this framework's tests and benches
MODEL_DIMS = {
    "vision": {"clip": 768, "dino": 768, "resnet": 2048, "convnext": 1024},
    "language": {"sentence-bert": 384, "mpnet": 768, "bert": 768, "roberta": 768},
}

_IGNORED_PREFIXES = ("vision_model.", "language_model.", "clip_text_model.", "vision_contrastive_projection.",
                     "text_contrastive_projection.", "temperature")


def _resolve_dim(kind: str, name: Optional[str], explicit: Optional[int]) -> int:
    if explicit is not None:
        return int(explicit)
    if not name:
        return 0
    if name in MODEL_DIMS[kind]:
        return MODEL_DIMS[kind][name]
    if name.startswith("cached") and name[6:].isdigit():
        return int(name[6:])
    raise ValueError(f"{kind.capitalize()} model '{name}' not found in MODEL_CONFIGS.")


def _activation_module(name: str) -> nn.Module:
    return {"relu": nn.ReLU(), "gelu": nn.GELU(), "tanh": nn.Tanh(), "leaky_relu": nn.LeakyReLU(),
            "silu": nn.SiLU()}.get((name or "relu").lower(), nn.ReLU())


class _GatedFusionParams(nn.Module):
    """Parameter container with the key names of GatedFusionLayer (layers.py:189-192)."""

    def __init__(self, D, M):
        super().__init__()
        self.gating_network = nn.Sequential(nn.Linear(D * M, M), nn.Softmax(dim=-1))


class _AttentionFusionParams(nn.Module):
    """Parameter container with the key names of AttentionFusionLayer (layers.py:124-131)."""

    def __init__(self, D, heads, dropout):
        super().__init__()
        self.attention = nn.MultiheadAttention(embed_dim=D, num_heads=heads, dropout=dropout, batch_first=False)
        self.norm = nn.LayerNorm(D)


class FastMultimodalRecommender(nn.Module):
    def __init__(self, n_users: int, n_items: int, n_tags: int, num_numerical_features: int,
                 embedding_dim: int = 128, vision_model_name: Optional[str] = "clip",
                 language_model_name: Optional[str] = "sentence-bert", freeze_vision: bool = True,
                 freeze_language: bool = True, use_contrastive: bool = True, dropout_rate: float = 0.3,
                 num_attention_heads: int = 4, attention_dropout: float = 0.1,
                 fusion_hidden_dims: List[int] = None, fusion_activation: str = "relu",
                 use_batch_norm: bool = True, projection_hidden_dim: Optional[int] = None,
                 final_activation: str = "sigmoid", init_method: str = "xavier_uniform",
                 contrastive_temperature: float = 0.07, fusion_type: str = "concatenate",
                 vision_dim: Optional[int] = None, language_dim: Optional[int] = None, kernel_path: str = "auto",
                 operand_dtype: str = "bf16", exact_rescore: bool = True):
        super().__init__()
        self.fusion_type = fusion_type
        self.n_users, self.n_items, self.n_tags = n_users, n_items, n_tags
        self.embedding_dim = embedding_dim
        self.num_numerical_features = num_numerical_features
        self.vision_model_name, self.language_model_name = vision_model_name, language_model_name
        self.vision_dim = _resolve_dim("vision", vision_model_name, vision_dim)
        self.language_dim = _resolve_dim("language", language_model_name, language_dim)
        self.num_attention_heads = num_attention_heads
        self.fusion_hidden_dims = list(fusion_hidden_dims or [512, 256, 128])
        self.fusion_activation = fusion_activation
        self.use_batch_norm = use_batch_norm
        self.projection_hidden_dim = projection_hidden_dim
        self.final_activation = final_activation
        self.kernel_path = kernel_path
        self.operand_dtype = operand_dtype      # 16-bit operand format of the fused tcgen05 kernels: "bf16" | "fp16"
        # exact mode of the fused path: the 64 candidates it keeps per user are re-scored in fp32 and re-ranked, so
        # full-catalogue lists carry the same fp32 scores as forward() / get_item_score (False: raw 16-bit scores)
        self.exact_rescore = bool(exact_rescore)
        # the contrastive heads and backbones are training / feature-production concerns (out of scope)
        self.use_contrastive = False
        self.vision_model = None
        self.language_model = None

        D = embedding_dim
        self.user_embedding = nn.Embedding(n_users, D)
        self.item_embedding = nn.Embedding(n_items, D)
        self.tag_embedding = nn.Embedding(n_tags, D)
        init = {"xavier_uniform": nn.init.xavier_uniform_, "xavier_normal": nn.init.xavier_normal_,
                "kaiming_uniform": lambda w: nn.init.kaiming_uniform_(w, nonlinearity="relu"),
                "kaiming_normal": lambda w: nn.init.kaiming_normal_(w, nonlinearity="relu")}.get(
                    (init_method or "").lower(), nn.init.xavier_uniform_)
        for emb in (self.user_embedding, self.item_embedding, self.tag_embedding):
            init(emb.weight)

        def projection(in_dim):
            act = _activation_module(fusion_activation)
            if projection_hidden_dim:
                return nn.Sequential(nn.Linear(in_dim, projection_hidden_dim), act, nn.Dropout(dropout_rate),
                                     nn.Linear(projection_hidden_dim, D), act, nn.Dropout(dropout_rate))
            return nn.Sequential(nn.Linear(in_dim, D), act, nn.Dropout(dropout_rate))

        if self.vision_dim:
            self.vision_projection = projection(self.vision_dim)
        if self.language_dim:
            self.language_projection = projection(self.language_dim)
        self.numerical_projection = projection(num_numerical_features) if num_numerical_features > 0 else None

        M = 3 + (self.vision_dim > 0) + (self.language_dim > 0) + (num_numerical_features > 0)
        self.num_modalities = M
        self.fusion_layer = None
        if fusion_type == "concatenate":
            fusion_in = M * D
        elif fusion_type == "attention":
            self.fusion_layer = _AttentionFusionParams(D, num_attention_heads, attention_dropout)
            fusion_in = D
        elif fusion_type == "gated":
            self.fusion_layer = _GatedFusionParams(D, M)
            fusion_in = D
        else:
            raise ValueError(f"Unknown fusion type: '{fusion_type}'")

        layers: List[nn.Module] = []
        act = _activation_module(fusion_activation)
        in_dim = fusion_in
        for hdim in self.fusion_hidden_dims:
            layers.append(nn.Linear(in_dim, hdim))
            layers.append(act)
            if use_batch_norm:
                layers.append(nn.BatchNorm1d(hdim))
            layers.append(nn.Dropout(dropout_rate))
            in_dim = hdim
        layers.append(nn.Linear(in_dim, 1))
        if final_activation == "sigmoid":
            layers.append(nn.Sigmoid())
        elif final_activation == "tanh":
            layers.append(nn.Tanh())
        self.prediction_network = nn.Sequential(*layers)

        # role -> [engine, weights version]; "forward" holds per-call item rows, "catalogue"
        # the recommender's resident item records, so forward() never clobbers a catalogue
        self._engines: Dict[str, list] = {}
        self.eval()

    # ------------------------------------------------------------ checkpoint
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts a reference checkpoint unchanged: backbone and contrastive-head
        entries are dropped, everything else must match (strict)."""
        filtered = {k: v for k, v in state_dict.items() if not k.startswith(_IGNORED_PREFIXES)}
        out = super().load_state_dict(filtered, strict=strict, assign=assign)
        for slot in self._engines.values():
            slot[1] = None
        return out

    # ---------------------------------------------------------------- engine
    def _weights_version(self):
        dev = self.user_embedding.weight.device
        return (str(dev),) + tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def engine(self, role: str = "catalogue") -> PxrEngine:
        """The native handle of ``role`` on the model's current device, (re)loaded
        when a parameter tensor changed."""
        dev = self.user_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("FastMultimodalRecommender runs on CUDA only (no CPU fallback); call .to('cuda')")
        ver = self._weights_version()
        slot = self._engines.setdefault(role, [None, None])
        if slot[0] is None or slot[0].device != dev:
            slot[0] = PxrEngine(
                fusion_type=self.fusion_type, embedding_dim=self.embedding_dim, vision_dim=self.vision_dim,
                language_dim=self.language_dim, num_numerical=self.num_numerical_features,
                hidden_dims=self.fusion_hidden_dims, n_tags=self.n_tags, num_heads=self.num_attention_heads,
                activation=self.fusion_activation, final_activation=self.final_activation,
                use_batch_norm=self.use_batch_norm, projection_hidden_dim=self.projection_hidden_dim,
                path=self.kernel_path, precision=self.operand_dtype, device=dev, rescore=self.exact_rescore)
            slot[1] = None
        if slot[1] != ver:
            sd = {k: v for k, v in self.state_dict().items() if not k.endswith("num_batches_tracked")}
            eps = 1e-5
            for m in self.prediction_network:
                if isinstance(m, nn.BatchNorm1d):
                    eps = m.eps
            slot[0].load_weights(sd, self.use_batch_norm, eps)
            slot[1] = ver
        return slot[0]

    # --------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, user_idx: torch.Tensor, item_idx: torch.Tensor, tag_idx: torch.Tensor,
                image: Optional[torch.Tensor] = None, text_input_ids: Optional[torch.Tensor] = None,
                text_attention_mask: Optional[torch.Tensor] = None,
                numerical_features: Optional[torch.Tensor] = None,
                clip_text_input_ids: Optional[torch.Tensor] = None,
                clip_text_attention_mask: Optional[torch.Tensor] = None,
                return_embeddings: bool = False, debug_this_batch: bool = False, return_logits: bool = False):
        """(B,) indices + cached features -> (B, 1) fp32 scores, as reference
        ``forward`` (multimodal.py:528-610).  A modality whose tensor is None is
        an error here: the reference would silently change the fusion width."""
        if self.training:
            raise RuntimeError("FastMultimodalRecommender is inference-only: call .eval()")
        if return_embeddings:
            raise NotImplementedError("return_embeddings (contrastive training outputs) is out of scope")
        eng = self.engine("forward")
        dev = eng.device
        B = int(user_idx.shape[0])
        if image is not None and image.dim() != 2:
            raise ValueError("image must be the cached (B, vision_dim) backbone feature, not raw pixels: "
                             "the frozen backbones are hoisted out of the scoring path")
        if text_input_ids is not None and (text_input_ids.dim() != 2 or not text_input_ids.is_floating_point()):
            raise ValueError("text_input_ids must be the cached (B, language_dim) float text feature")
        if self.vision_dim and image is None:
            raise ValueError("vision features are required by this model configuration")
        if self.language_dim and (text_input_ids is None or text_attention_mask is None):
            raise ValueError("text features (and a non-None attention mask) are required by this configuration")
        if self.num_numerical_features and numerical_features is None:
            raise ValueError("numerical_features are required by this model configuration")
        if B == 0:
            return torch.empty((0, 1), dtype=torch.float32, device=dev)
        check_index_range(user_idx, self.n_users, "user_idx")          # nn.Embedding raises IndexError (multimodal.py:553)
        eng.precompute_items(self.item_embedding.weight, tag_idx, image, text_input_ids, numerical_features,
                             item_idx=item_idx, n_rows=B)                   # checks item_idx / tag_idx ranges
        rows = torch.arange(B, dtype=torch.int64, device=dev)
        uidx = user_idx.to(device=dev, dtype=torch.int64).contiguous()
        uemb = self.user_embedding.weight.detach()
        if return_logits:
            s, z = eng.score_pairs(uemb, uidx, rows, want_logit=True)
            return s.unsqueeze(1), z.unsqueeze(1)
        return eng.score_pairs(uemb, uidx, rows).unsqueeze(1)


# the reference exposes this alias too (multimodal.py, "Backward compatibility alias")
PretrainedMultimodalRecommender = FastMultimodalRecommender
