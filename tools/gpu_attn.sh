#!/bin/bash
WF_ATTN_H8=2 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 2 1 2; do echo "h8=$v"; WF_ATTN_H8=$v bash tools/gpu_quick.sh 2>&1 | python -c "
import sys,re,ast
t=sys.stdin.read(); m=re.search(r'\{.*\}', t, re.S); d=ast.literal_eval(m.group(0)); print(t.splitlines()[0]); print({k:d[k] for k in ['attn_fwd','attn_fwd_stats','attn_bwd','attn_bwd_stats']})"; done
