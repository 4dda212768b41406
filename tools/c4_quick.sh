for P in primary shadow fused frame; do timeout 120 python tools/prof_configs.py c4 $P -1 8 | tail -1; done
timeout 120 python tools/prof_configs.py c2 primary -1 8 | tail -1
