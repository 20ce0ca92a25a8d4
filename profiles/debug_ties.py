import sys, os, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gpu_harness as h
syn = h.PKG.synthetic
D, N, T, VOCAB, K, MLP = 512, 196, 26, 10000, 1001, 1024
def conv_params(net):
    pc = net.question_encoder.phrase_conv_pool
    u, b, t = pc.conv_unigram[1], pc.conv_bigram[1], pc.conv_trigram[1]
    return [q.detach() for q in (u.weight, u.bias, b.weight, b.bias, t.weight, t.bias)]
for pdl in ("1", "0"):
    h.PKG._lib.set_option("pdl", pdl)
    for B, sw in ((24, 1.0), (12, 8.0), (24, 1.0), (12, 8.0), (160, 1.0)):
        p = syn.make_params(D, VOCAB, K, MLP, seed=3)
        pre = "question_encoder.phrase_conv_pool."
        for n in ("unigram", "bigram", "trigram"):
            p[pre + f"conv_{n}.1.weight"] = p[pre + f"conv_{n}.1.weight"] * np.float32(sw)
        x = syn.make_inputs(B, N, T, D, VOCAB, K, seed=9, min_len=1)
        net = h.build_net(p, D, VOCAB, K, MLP)
        word = torch.from_numpy(p["question_encoder.word_embedding.weight"][x["tokens"]]).cuda()
        lens = torch.from_numpy(x["lens"]).cuda()
        for rep in range(3):
            with torch.no_grad():
                out, idx, saved = h.PKG.ops.phrase_conv_pool(word, *conv_params(net), lens)
            st = saved[-256:-224].view(torch.int32).cpu()
            f = saved[-248:-232].view(torch.float32).cpu()
            wn0 = float(np.sqrt((p[pre + "conv_unigram.1.weight"][0].astype(np.float64) ** 2).sum()))
            xn0 = float((word[0, 0].double() ** 2).sum())
            print(f"pdl={pdl} B={B} w x{sw} rep{rep}: found {int(st[0])} cap {int(st[1])} xn2[0] {float(f[0]):.4f} (want {xn0:.4f}) wn[0] {float(f[1]):.4f} (want {wn0:.4f}) wn[last] {float(f[2]):.4f} xn2[last] {float(f[3]):.4f}", flush=True)
