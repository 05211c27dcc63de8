# all-pairs diagnostics: B200CORR_DEBUG bits (1 = no level-0 stores, 2 = no pooled stores) x CTA-pair / single-CTA kernel
for c in 2 1; do for d in 0 1 3; do B200CORR_ALLPAIRS_CTAS=$c B200CORR_DEBUG=$d timeout 60 python scripts/raft_debug_timing.py | sed "s/^/ctas $c: /"; done; done
