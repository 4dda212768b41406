python tools/prof_configs.py c4 primary -1 8
python tools/prof_configs.py c4 primary 1 8
python tools/prof_configs.py c4 shadow -1 8
python tools/prof_configs.py c4 shadow 1 8
python tools/prof_configs.py c4 fused -1 8
python tools/prof_configs.py c4 primary -1 8 tile_hints=0
python tools/prof_configs.py c4 fused -1 8 tile_hints=0
python tools/prof_configs.py c4 primary -1 8 top_pairs=524287 l2_persist_kb=32768
python tools/prof_configs.py c4 shadow -1 8 top_pairs=524287 l2_persist_kb=32768
python tools/prof_configs.py c4 diffuse -1 8 top_pairs=524287 l2_persist_kb=32768
python tools/prof_configs.py c4 diffuse -1 8
python tools/prof_configs.py c4 primary -1 8 top_pairs=524287
python tools/prof_configs.py c1 primary -1 8
python tools/prof_configs.py c1 primary 1 8
python tools/prof_configs.py c1 fused -1 8
