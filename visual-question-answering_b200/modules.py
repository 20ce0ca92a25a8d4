"""nn.Module mirrors of the reference's Hierarchical Co-Attention model (reference model.py:157-434).

Same class names, constructor arguments, ``forward`` signatures, sub-module attribute names and
``state_dict`` keys as the reference, so ``from model import HierarchicalCoAttentionNet`` in the
reference's main.py (main.py:15,164) picks these up unchanged and checkpoints interchange
(main.py:168-176, 260-263).  The arithmetic of the hot path runs in the hand-written sm_100a kernels
behind the ``hiecoattn::*`` custom ops (ops.py); parameters are held by stock ``nn.Linear`` /
``nn.Conv1d`` / ``nn.Embedding`` containers only so that names, shapes and default initialisation match.

Out of the hot path and therefore stock PyTorch here: the VGG11-bn trunk (torchvision) and the GRU baseline
network.  The sentence ``nn.LSTM`` container only holds the weights: the recurrence itself runs in the
persistent kernel of csrc/lstm.cu (cuDNN is the fallback for hidden sizes that kernel does not cover).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from . import ops


class QuestionLens:
    """Sequence lengths held on both sides of the PCIe bus.

    The reference moves ``ques_len`` to the GPU (main.py:207) and then hands it to
    ``pack_padded_sequence`` (model.py:287), which needs it on the CPU: a device->host sync per step (and
    an error on current PyTorch).  Passing a ``QuestionLens`` instead of a tensor gives the encoder both
    copies up front, so the step stays free of host synchronisation and can be captured in a CUDA graph.
    A plain tensor (CPU or CUDA) is still accepted everywhere.
    """

    def __init__(self, lens_cpu: Tensor, device=None, lens_dev: Optional[Tensor] = None):
        self.cpu = lens_cpu.detach().to("cpu", torch.int64)
        self.dev = lens_dev if lens_dev is not None else self.cpu.to(device, non_blocking=True)


class _cudnn_fp32:
    """cuDNN RNNs default to TF32 math (torch.backends.cudnn.allow_tf32 = True), which alone costs ~1e-3 of
    relative error in the sentence features and their gradients -- the whole north_star budget.  The
    sentence LSTM therefore runs with TF32 disabled, i.e. in the reference's fp32."""

    def __enter__(self):
        self.prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cudnn.allow_tf32 = self.prev
        return False


def _check_lens(lens_cpu: Tensor, max_len: int) -> None:
    """The misuse pack_padded_sequence rejects in the reference (model.py:287): zero / negative or unsorted lengths."""
    if lens_cpu.numel() and int(lens_cpu.min()) <= 0:
        raise RuntimeError("Length of all samples has to be greater than 0, but found an element in 'lengths' that is <= 0")
    if lens_cpu.numel() > 1 and bool((lens_cpu[1:] > lens_cpu[:-1]).any()):
        raise RuntimeError("`lengths` array must be sorted in decreasing order when `enforce_sorted` is True.")
    if lens_cpu.numel() and int(lens_cpu.max()) > max_len:
        raise RuntimeError(f"a sequence length ({int(lens_cpu.max())}) exceeds the padded length {max_len}")


def _f32(t: Tensor) -> Tensor:
    return t if t.dtype == torch.float32 else t.float()


class PhraseConvPool(nn.Module):
    """Unigram / bigram / trigram Conv1d + tanh, max over consecutive channel triples (model.py:301-334)."""

    def __init__(self, emb_dim):
        super().__init__()
        # index 1 inside each Sequential keeps the reference's state_dict keys ("conv_unigram.1.weight", ...)
        self.conv_unigram = nn.Sequential(nn.ConstantPad1d((0, 0), 0), nn.Conv1d(emb_dim, emb_dim, 1, 1), nn.Tanh())
        self.conv_bigram = nn.Sequential(nn.ConstantPad1d((1, 0), 0), nn.Conv1d(emb_dim, emb_dim, 2, 1), nn.Tanh())
        self.conv_trigram = nn.Sequential(nn.ConstantPad1d((1, 1), 0), nn.Conv1d(emb_dim, emb_dim, 3, 1), nn.Tanh())
        self.max_pool = nn.MaxPool2d(kernel_size=(1, 3))        # parameter-free; kept for attribute parity

    def forward(self, x_question: Tensor, lens_dev: Optional[Tensor] = None, return_indices: bool = False):
        """x_question [B,T,E] -> [B,T,E].  ``lens_dev`` (optional, int64 on the GPU) zeroes rows t >= len."""
        u, b, t = self.conv_unigram[1], self.conv_bigram[1], self.conv_trigram[1]
        out, idx, _ = ops.phrase_conv_pool(_f32(x_question), u.weight, u.bias, b.weight, b.bias, t.weight, t.bias, lens_dev)
        return (out, idx) if return_indices else out


class QuestionCoAttentionEncoder(nn.Module):
    """Word embedding -> PhraseConvPool -> LSTM; returns (word, phrase, sentence) [B,T,d] (model.py:246-298)."""

    def __init__(self, vocab_size, word_emb_dim, hidden_dim):
        super().__init__()
        self.vocab_size = vocab_size
        self.embedding_dim = word_emb_dim
        self.hidden_dim = hidden_dim
        self.word_embedding = nn.Embedding(self.vocab_size, self.embedding_dim, padding_idx=0)
        self.phrase_conv_pool = PhraseConvPool(self.embedding_dim)
        self.sentence_lstm = nn.LSTM(self.embedding_dim, self.hidden_dim)

    def forward(self, x: Tensor, x_lens) -> Tuple[Tensor, Tensor, Tensor]:
        max_seq_len = x.shape[1]
        if isinstance(x_lens, QuestionLens):
            lens_cpu, lens_dev = x_lens.cpu, x_lens.dev
        else:
            lens_cpu = x_lens.cpu() if x_lens.is_cuda else x_lens          # the one D2H sync the reference also pays
            lens_dev = x_lens
        # the kernels read the lengths as int64 on the tokens' device, whatever the caller's loader produced (int32, CPU, ...)
        lens_cpu = lens_cpu.to(torch.int64)
        lens_dev = lens_dev.to(x.device, torch.int64)
        x_word_emb = ops.embedding(x, self.word_embedding.weight)                       # model.py:282
        # phrase level with the pad rows already zeroed (what pack -> pad does at model.py:287,292)
        x_phrase_emb = self.phrase_conv_pool(x_word_emb, lens_dev)                      # model.py:284
        _check_lens(lens_cpu, max_seq_len)                                              # what pack_padded_sequence raises on
        lstm = self.sentence_lstm
        B, T, E = x_phrase_emb.shape
        if ops.lstm_supported(B, T, E, lstm.hidden_size) and lstm.num_layers == 1 and not lstm.bidirectional:
            # pack -> LSTM -> pad (model.py:287-296) as one persistent tcgen05 recurrence over the padded batch (csrc/lstm.cu)
            x_sentence_emb = ops.lstm(x_phrase_emb, lens_dev, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0)[0]
        else:
            # widths the kernel does not cover: stock cuDNN over the padded batch.  The LSTM is causal, so running it over the
            # zero-padded batch and zeroing rows t >= len afterwards gives the packed run's outputs (and, through the mask, its
            # gradients) without the per-timestep gather / scatter copies of pack / pad
            with _cudnn_fp32():
                x_sentence_emb, _ = lstm(x_phrase_emb.transpose(0, 1))                    # model.py:289
            valid = torch.arange(max_seq_len, device=x.device)[:, None] < lens_dev[None, :]
            x_sentence_emb = (x_sentence_emb * valid[..., None]).transpose(0, 1)
        return x_word_emb, x_phrase_emb, x_sentence_emb


class ParallelCoAttention(nn.Module):
    """Parallel co-attention applied at the word, phrase and sentence level (model.py:337-397)."""

    def __init__(self, hidden_dim):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.W_b = nn.Linear(self.hidden_dim, self.hidden_dim)   # declared, never used by the reference (model.py:347,377)
        self.W_v = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.W_q = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.w_v = nn.Linear(self.hidden_dim, 1)
        self.w_q = nn.Linear(self.hidden_dim, 1)

    def forward_stacked(self, x_img: Tensor, x_ques_hierarchy: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
        """-> (vhat [3,B,d], qhat [3,B,d])"""
        q0, q1, q2 = (_f32(q) for q in x_ques_hierarchy)
        out = ops.coattn(_f32(x_img), q0, q1, q2, self.W_v.weight, self.W_v.bias, self.W_q.weight, self.W_q.bias,
                         self.w_v.weight, self.w_v.bias, self.w_q.weight, self.w_q.bias)
        return out[0], out[1]

    @torch.no_grad()
    def attention_maps(self, x_img: Tensor, x_ques_hierarchy: Sequence[Tensor]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """Inference-time export of the co-attention: (vhat [3,B,d], qhat [3,B,d], a_v [B,3,N], a_q [B,3,T]) -- the weights over
        the N image regions and over the T token positions at the word / phrase / sentence level (model.py:387-388).  The
        reference computes them and throws them away; its README leaves visualisation as a TO-DO."""
        q0, q1, q2 = (_f32(q) for q in x_ques_hierarchy)
        vhat, qhat, saved = ops.coattn(_f32(x_img), q0, q1, q2, self.W_v.weight, self.W_v.bias, self.W_q.weight, self.W_q.bias,
                                       self.w_v.weight, self.w_v.bias, self.w_q.weight, self.w_q.bias)
        B, N, d = x_img.shape
        a_v, a_q = ops.coattn_attention_maps(saved, B, N, q0.shape[1], d)
        return vhat, qhat, a_v.clone(), a_q.clone()

    def forward(self, x_img: Tensor, x_ques_hierarchy: Sequence[Tensor]) -> Tuple[List[Tensor], List[Tensor]]:
        if len(x_ques_hierarchy) != 3:
            raise ValueError("ParallelCoAttention expects the (word, phrase, sentence) hierarchy: 3 tensors")
        vhat, qhat = self.forward_stacked(x_img, x_ques_hierarchy)
        return list(vhat.unbind(0)), list(qhat.unbind(0))


class MLPClassifier(nn.Module):
    """Recursive MLP over the three levels of attended features (model.py:400-434)."""

    def __init__(self, hidden_dim, mlp_dim, K):
        super().__init__()
        self.W_w = nn.Linear(hidden_dim, hidden_dim)
        self.W_p = nn.Linear(2 * hidden_dim, hidden_dim)
        self.W_s = nn.Linear(2 * hidden_dim, mlp_dim)
        self.W_h = nn.Linear(mlp_dim, K)

    def forward_stacked(self, vhat: Tensor, qhat: Tensor) -> Tensor:
        return ops.mlp(vhat, qhat, self.W_w.weight, self.W_w.bias, self.W_p.weight, self.W_p.bias, self.W_s.weight,
                       self.W_s.bias, self.W_h.weight, self.W_h.bias)[0]

    def forward(self, x_img_feats: Sequence[Tensor], x_ques_feats: Sequence[Tensor]) -> Tensor:
        vhat = torch.stack([_f32(t) for t in x_img_feats])
        qhat = torch.stack([_f32(t) for t in x_ques_feats])
        return self.forward_stacked(vhat, qhat)


class ImageCoAttentionEncoder(nn.Module):
    """VGG11-bn trunk, 448x448 -> [B, 196, 512] (model.py:190-243).  Stock torchvision: out of the hot path."""

    def __init__(self, is_trainable, weights_path):
        super().__init__()
        self.is_trainable = is_trainable
        self.weights_path = weights_path
        self.vgg11_encoder = self.build_vgg_encoder()
        self.flatten = nn.Flatten(start_dim=2, end_dim=3)

    def forward(self, x_img):
        return self.flatten(self.vgg11_encoder(x_img)).permute(0, 2, 1)      # non-contiguous [B, H*W, 512] view

    def build_vgg_encoder(self):
        import torchvision.models as models
        vgg11 = models.vgg11_bn(weights=None if self.weights_path else "IMAGENET1K_V1")
        if self.weights_path:
            vgg11.load_state_dict(torch.load(self.weights_path))
        enc = vgg11.features
        if not self.is_trainable:
            for p in enc.parameters():
                p.requires_grad = False
        return enc


class HierarchicalCoAttentionNet(nn.Module):
    """Drop-in for the reference's ``--model attention`` network (model.py:157-187)."""

    def __init__(self, ques_enc_params, img_enc_params, K, mlp_dim=1024):
        super().__init__()
        self.hidden_dim = ques_enc_params["hidden_dim"]
        self.image_encoder = ImageCoAttentionEncoder(**img_enc_params)
        self.question_encoder = QuestionCoAttentionEncoder(**ques_enc_params)
        self.co_attention = ParallelCoAttention(self.hidden_dim)
        self.mlp_classify = MLPClassifier(self.hidden_dim, mlp_dim, K)

    def forward(self, x_img, x_ques, x_ques_lens):
        x_word, x_phrase, x_sentence = self.question_encoder(x_ques, x_ques_lens)
        x_img_features = self.image_encoder(x_img)
        return self.forward_features(x_img_features, (x_word, x_phrase, x_sentence))

    def forward_features(self, x_img_features, x_ques_features):
        """Co-attention + classifier on precomputed [B,N,d] image features (synthetic-feature entry point)."""
        vhat, qhat = self.co_attention.forward_stacked(x_img_features, x_ques_features)
        return self.mlp_classify.forward_stacked(vhat, qhat)


class CrossEntropyLoss(nn.Module):
    """``nn.CrossEntropyLoss()`` as the reference's loop builds it (main.py:179: default mean reduction, no weights) with the
    loss and its gradient computed in one launch; ``scale`` folds a constant factor (1 / world size in data-parallel runs) in."""

    def __init__(self, scale: float = 1.0):
        super().__init__()
        self.scale = float(scale)

    def forward(self, logits: Tensor, labels: Tensor) -> Tensor:
        return ops.cross_entropy(logits, labels, self.scale)


class HieCoAttnHotPath(nn.Module):
    """question encoder -> co-attention x3 -> MLP on precomputed image features.

    This is the path BASELINE.json benchmarks (the VGG trunk is bypassed with synthetic 196x512 grids);
    it shares its three sub-modules' names with HierarchicalCoAttentionNet so state_dicts interchange."""

    def __init__(self, vocab_size=10000, hidden_dim=512, K=1001, mlp_dim=1024):
        super().__init__()
        self.question_encoder = QuestionCoAttentionEncoder(vocab_size, hidden_dim, hidden_dim)
        self.co_attention = ParallelCoAttention(hidden_dim)
        self.mlp_classify = MLPClassifier(hidden_dim, mlp_dim, K)

    def forward(self, x_img_features, x_ques, x_ques_lens):
        hier = self.question_encoder(x_ques, x_ques_lens)
        vhat, qhat = self.co_attention.forward_stacked(x_img_features, hier)
        return self.mlp_classify.forward_stacked(vhat, qhat)

    @torch.no_grad()
    def predict(self, x_img_features, x_ques, x_ques_lens, topk: int = 5, return_attention: bool = False):
        """Inference (the reference's unimplemented ``--mode test``, main.py:286-287): top-k answer classes with their softmax
        probabilities, optionally with the attention maps (a_v [B,3,N] over regions, a_q [B,3,T] over tokens)."""
        hier = self.question_encoder(x_ques, x_ques_lens)
        vhat, qhat, a_v, a_q = self.co_attention.attention_maps(x_img_features, hier)
        logits = self.mlp_classify.forward_stacked(vhat, qhat)
        prob, idx = torch.softmax(logits, dim=1).topk(min(topk, logits.shape[1]), dim=1)
        return (prob, idx, a_v, a_q) if return_attention else (prob, idx)


class InferenceSession:
    """The validation / serving call of the reference (main.py:301-335: ``eval()``, ``no_grad()``, ``model(image, question, ques_len)``
    per batch) for ONE fixed shape, replayed from a CUDA graph: the ~40 kernel launches of a forward cost more host time than device
    time at small batches (0.56 ms eager at batch 1), a replay costs one launch.

    ``session = InferenceSession(net, batch, regions, tokens)`` captures ``net(feats, tokens, lens)`` on static buffers;
    ``session(feats, tokens, lens)`` copies the inputs in (device or host tensors), replays, and returns the static ``[batch, K]`` logits
    (clone them to keep them across calls).  Works for ``HieCoAttnHotPath`` (feature grids) -- the VGG trunk of the full wrapper is
    stock cuDNN and can be captured the same way by passing images of the capture-time size."""

    def __init__(self, net: nn.Module, batch: int, regions: int, tokens: int, device=None, feature_dim: Optional[int] = None):
        self.net = net.eval()
        dev = torch.device(device) if device is not None else next(net.parameters()).device
        d = feature_dim or net.co_attention.hidden_dim
        self.feats = torch.zeros(batch, regions, d, device=dev)
        self.tokens = torch.ones(batch, tokens, dtype=torch.int64, device=dev)
        self.lens_dev = torch.full((batch,), tokens, dtype=torch.int64, device=dev)
        self._lens = QuestionLens(torch.full((batch,), tokens, dtype=torch.int64), dev, self.lens_dev)
        self.max_len = tokens
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                                  # warm-up outside the capture (lazy initialisation, allocator)
                self.net(self.feats, self.tokens, self._lens)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.logits = self.net(self.feats, self.tokens, self._lens)

    @torch.no_grad()
    def __call__(self, feats: Tensor, tokens: Tensor, lens) -> Tensor:
        lens_cpu = lens.cpu if isinstance(lens, QuestionLens) else (lens.cpu() if lens.is_cuda else lens)
        _check_lens(lens_cpu.to(torch.int64), self.max_len)       # what pack_padded_sequence raises on (model.py:287), checked on the host copy
        self.feats.copy_(feats, non_blocking=True)
        self.tokens.copy_(tokens, non_blocking=True)
        self.lens_dev.copy_(lens.dev if isinstance(lens, QuestionLens) else lens, non_blocking=True)
        self.graph.replay()
        return self.logits


# ---------------------------------------------------------------------------------------------------------------
# Baseline network (reference model.py:10-151).  Not on the accelerated path (BASELINE.json config 2 only
# times it for context); stock PyTorch layers with the reference's names so `from model import VQABaselineNet`
# keeps working.


class ImageBaselineEncoder(nn.Module):
    def __init__(self, is_trainable, weights_path):
        super().__init__()
        self.is_trainable = is_trainable
        self.weights_path = weights_path
        self.vgg11_encoder = self.build_vgg_encoder()
        self.embedding_layer = nn.Sequential(nn.Linear(4096, 1024), nn.Tanh())

    def build_vgg_encoder(self):
        import torchvision.models as models
        vgg11 = models.vgg11_bn(weights=None if self.weights_path else "IMAGENET1K_V1")
        if self.weights_path:
            vgg11.load_state_dict(torch.load(self.weights_path))
        fc = nn.Sequential(nn.Flatten(), *list(vgg11.classifier)[:-1])
        enc = nn.Sequential(OrderedDict(conv_layers=vgg11.features, avgpool=vgg11.avgpool, fc_layers=fc))
        if not self.is_trainable:
            for p in enc.parameters():
                p.requires_grad = False
        return enc

    def forward(self, x_img):
        return self.embedding_layer(F.normalize(self.vgg11_encoder(x_img), dim=1, p=2))


class QuestionBaselineEncoder(nn.Module):
    def __init__(self, vocab_size, word_emb_dim, hidden_dim):
        super().__init__()
        self.hidden_dim, self.vocab_size, self.word_emb_dim = hidden_dim, vocab_size, word_emb_dim
        self.word_embedding = nn.Sequential(nn.Embedding(vocab_size, word_emb_dim), nn.Tanh())
        self.gru = nn.GRU(word_emb_dim, hidden_dim)
        self.embedding_layer = nn.Sequential(nn.Linear(hidden_dim, 1024), nn.Tanh())

    def forward(self, x, seq_lengths):
        if isinstance(seq_lengths, QuestionLens):
            seq_lengths = seq_lengths.cpu
        elif seq_lengths.is_cuda:
            seq_lengths = seq_lengths.cpu()
        packed = pack_padded_sequence(self.word_embedding(x), seq_lengths, batch_first=True)
        _, hidden = self.gru(packed)
        return self.embedding_layer(hidden.squeeze(0))


class VQABaselineNet(nn.Module):
    def __init__(self, ques_enc_params, img_enc_params, K):
        super().__init__()
        self.image_encoder = ImageBaselineEncoder(**img_enc_params)
        self.question_encoder = QuestionBaselineEncoder(**ques_enc_params)
        self.mlp = nn.Sequential(nn.Linear(1024, 1000), nn.Dropout(0.5), nn.Tanh())
        self.fc_final = nn.Linear(1000, K)

    def forward(self, x_img, x_ques, x_ques_len):
        return self.fc_final(self.mlp(self.image_encoder(x_img) * self.question_encoder(x_ques, x_ques_len)))
