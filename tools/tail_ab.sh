# development aid: k_tail hand-over threshold sweep (RL_TAIL_MAX), full frame share and one rank's share of an 8-GPU frame
for t in 0 37888 75776 151552 227328 303104; do RL_TAIL_MAX=$t python tools/ab_variants.py rustlight_b200/librl_b200.so | sed "s/^/TAIL=$t /"; done
for t in 0 75776 151552 303104; do RL_TAIL_MAX=$t python tools/ab_rank.py 2>&1 | tail -1 | sed "s/^/TAIL=$t /"; done
