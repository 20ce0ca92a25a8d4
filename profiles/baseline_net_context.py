"""BASELINE.json configs[1] (context number only, not on the accelerated path): the reference's `VQABaselineNet` (model.py:10-151: VGG11-bn ->
4096 -> L2 norm -> Linear + tanh; Embedding(300) + tanh -> GRU(1024) -> Linear + tanh; product; MLP) as stock PyTorch modules on one B200,
batch 160, K = 1001, random-init VGG from a generated weights file, images randn[160, 3, 224, 224]; fwd + CE + bwd + Adam; and the
post-VGG remainder alone on synthetic 4096-d features."""
import importlib, json, os, sys, tempfile, torch, torchvision
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("visual-question-answering_b200")
dev = torch.device("cuda:0")
B, K, vocab, T = 160, 1001, 10000, 26
path = os.path.join(tempfile.mkdtemp(), "vgg.pth")
torch.manual_seed(0)
torch.save(torchvision.models.vgg11_bn(weights=None).state_dict(), path)
net = pkg.VQABaselineNet(dict(vocab_size=vocab, word_emb_dim=300, hidden_dim=1024), dict(is_trainable=False, weights_path=path), K=K).to(dev)
opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], 1e-4)
g = torch.Generator().manual_seed(1)
img = torch.randn(B, 3, 224, 224, generator=g).to(dev)
lens = torch.randint(3, T + 1, (B,), generator=g).sort(descending=True).values
tok = torch.zeros(B, T, dtype=torch.long)
for b in range(B):
    tok[b, :lens[b]] = torch.randint(1, vocab, (int(lens[b]),), generator=g)
tok, lab = tok.to(dev), torch.randint(0, K, (B,), generator=g).to(dev)


def timed(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def full_step():
    loss = torch.nn.functional.cross_entropy(net(img, tok, lens), lab)
    opt.zero_grad()
    loss.backward()
    opt.step()


feat4096 = torch.randn(B, 4096, generator=g).to(dev)


def tail_step():      # everything after the VGG trunk, on synthetic 4096-d features
    x_img = net.image_encoder.embedding_layer(torch.nn.functional.normalize(feat4096, dim=1, p=2))
    loss = torch.nn.functional.cross_entropy(net.fc_final(net.mlp(x_img * net.question_encoder(tok, lens))), lab)
    opt.zero_grad()
    loss.backward()
    opt.step()


out = {"what": "VQABaselineNet (reference model.py:10-151, stock PyTorch) on one B200, batch 160, K = 1001, fwd + CE + bwd + Adam, eager, fp32 (TF32 allowed by default in cuDNN convs)",
       "full_ms_per_step": timed(full_step), "post_vgg_ms_per_step": timed(tail_step)}
out["full_samples_per_s"] = B / out["full_ms_per_step"] * 1e3
out["post_vgg_samples_per_s"] = B / out["post_vgg_ms_per_step"] * 1e3
print(json.dumps(out))
