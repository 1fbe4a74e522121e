import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        print(line[:300]); continue
    d=json.loads(line)
    p=d.get('phases') or {}
    print("ms/step %.3f value %.0f e2e %s launches %s | esa %.2f (sort %.2f lcp %.2f ref %.2f cld %.2f tab %.2f) anchor %.2f (walk %.2f open %.2f br %.2f path %.2f asm %.2f filt %.2f) rows %.3f cmp %.3f | roof %s" % (
        d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('ms_per_step'), d.get('gpu_launches'),
        p.get('esa.total_ms',0),p.get('esa.sort_ms',0),p.get('esa.lcp_ms',0),p.get('esa.refine_ms',0),p.get('esa.cld_ms',0),p.get('esa.table_ms',0),
        p.get('anchor.total_ms',0),p.get('anchor.walk_ms',0),p.get('anchor.open_ms',0),p.get('anchor.bridge_ms',0),p.get('anchor.path_ms',0),p.get('anchor.assemble_ms',0),p.get('anchor.filter_ms',0),
        p.get('rows.ms',0),p.get('compare.ms',0), (d.get('roofline') or {}).get('frac')))
