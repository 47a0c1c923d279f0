"""Diagnostic: where does bf16 mode diverge from fp32 mode (== the reference) on one case?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import cases
from vltk_b200 import synthetic
from vltk_b200.frcnn import FRCNN
from vltk_b200.preprocess import Preprocess

def iou(a, b):
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2]-a[:, 0])*(a[:, 3]-a[:, 1]); ab = (b[:, 2]-b[:, 0])*(b[:, 3]-b[:, 1])
    return inter / np.maximum(aa[:, None] + ab[None, :] - inter, 1e-9)

case = sys.argv[1]
cfg, wseed, raws = cases.case_inputs(case)
sd = synthetic.make_state_dict(cfg, wseed)
ids, images, sizes, scales = Preprocess(cfg)(raws)
n = images.shape[0]; h4, w4 = cfg.res4_hw(images.shape[2], images.shape[3]); A = cfg.num_anchors
T = {}
for mode in ("fp32", "bf16"):
    env = os.environ.get("VLTK_NO_TC")
    m = FRCNN.from_pretrained(state_dict=sd, config=cfg, mode=mode)
    out = m(images, sizes, scales_yx=scales)
    T[mode] = dict(res4=m.debug_read("res4").reshape(n, h4, w4, -1), head=m.debug_read("rpn_head").reshape(n, h4*w4, -1),
                   topk=m.debug_read("topk_anchor_idx", np.int32).reshape(n, -1), props=m.debug_read("proposals").reshape(n, -1, 4),
                   plog=m.debug_read("proposal_logits").reshape(n, -1), cnt=m.debug_read("proposal_count", np.int32),
                   feats=m.debug_read("feats").reshape(n, -1, 2048), out=out)
    del m
a, b = T["fp32"], T["bf16"]
print(f"case {case}: res4 mean|err|/mean|x| {np.abs(a['res4']-b['res4']).mean()/np.abs(a['res4']).mean():.3e}")
hd, hb = a["head"], b["head"]
lg_a, lg_b = hd[..., 4*A:5*A], hb[..., 4*A:5*A]
dl_a, dl_b = hd[..., :4*A], hb[..., :4*A]
print(f"rpn logits: range [{lg_a.min():.2f},{lg_a.max():.2f}] std {lg_a.std():.3f}; abs err mean {np.abs(lg_a-lg_b).mean():.3e} max {np.abs(lg_a-lg_b).max():.3e}")
print(f"rpn deltas: std {dl_a.std():.4f}; abs err mean {np.abs(dl_a-dl_b).mean():.3e} max {np.abs(dl_a-dl_b).max():.3e}")
for i in range(n):
    sa, sb = set(a["topk"][i].tolist()), set(b["topk"][i].tolist())
    ca, cb = int(a["cnt"][i]), int(b["cnt"][i])
    # sorted logits spacing near the selected proposals
    la = np.sort(lg_a[i].reshape(-1))[::-1]
    print(f"img {i}: top-k set overlap {len(sa&sb)}/{len(sa)}; proposals {ca} vs {cb}; "
          f"logit of 1st/100th/300th/last-topk candidate {la[0]:.3f}/{la[99]:.3f}/{la[min(299,len(la)-1)]:.3f}/{la[len(sa)-1]:.3f}")
    pi = iou(b["props"][i, :cb], a["props"][i, :ca])
    print(f"       bf16 proposals re-found in fp32's list: IoU>=0.9 {(pi.max(1)>=0.9).mean():.2f}, IoU>=0.7 {(pi.max(1)>=0.7).mean():.2f}; "
          f"same rank & IoU>=0.9: {(np.diag(pi[:min(ca,cb),:min(ca,cb)])>=0.9).mean():.2f}")
    # how deep in the candidate list do the survivors sit?  (rank of the last kept proposal's logit)
    rank_last_a = int((la > a['plog'][i, ca-1]).sum()); 
    print(f"       fp32: last kept proposal is candidate #{rank_last_a} of {len(sa)}; logit gap between consecutive candidates around there ~{np.abs(np.diff(la[:rank_last_a+1])).mean():.2e}")
    j = pi.argmax(1); ok = pi.max(1) >= 0.9
    fa, fb = a["feats"][i, :ca][j[ok]], b["feats"][i, :cb][ok]
    cos = (fa*fb).sum(1)/(np.linalg.norm(fa,axis=1)*np.linalg.norm(fb,axis=1)+1e-12)
    print(f"       pooled-feature cosine of matched proposals: min {cos.min():.4f} mean {cos.mean():.4f} (n={ok.sum()})")
