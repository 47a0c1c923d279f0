for c in 8 4 1; do
  echo "== CHUNK=$c"
  VLTK_TCX_CHUNK=$c timeout 600 python tools/exact_probe.py golden 2>&1 | grep -v Warn | python -c "
import sys, json
for ln in sys.stdin:
    if ' {' not in ln: continue
    case, js = ln.split(' ', 1)
    try: d = json.loads(js)
    except Exception: continue
    e = d['exact_tc']
    print(case, 'ok' if all(e[k] for k in ('preds_per_image','obj_ids','attr_ids','n_props','topk_set')) else 'FAIL '+str({k:e[k] for k in ('preds_per_image','obj_ids','attr_ids','n_props','topk_set')}), 'res4 rms vs fp32', '%.2e'%d['res4_rel_rms_vs_fp32'], 'logit rms', '%.2e'%d['rpn_logit_rel_rms_vs_fp32'], 'feats', '%.2e'%d.get('feats_rel_rms_vs_fp32',-1), 'roi_feat vs golden', '%.2e'%e.get('roi_features_max_rel',-1), 'probs', '%.2e'%e.get('obj_probs_max_abs',-1))
"
done
