export RTB_PROBE_PARTS=1,8
for R in "3840 2160" "1920 1080"; do set -- $R
for P in "1 12 80 30 35" "1 30 80 30 50" "1 30 80 30 65" "1 45 80 30 50" "1 20 80 30 65" "1 12 80 30 65"; do set -- $R $P
echo "== $1x$2 hints=$3 heavy=$4 split=$5 keep=$6 light=$7"
RTB_TILE_HINTS=$3 RTB_HINT_HEAVY_PCT=$4 RTB_HINT_SPLIT_PCT=$5 RTB_HINT_KEEP_PCT=$6 RTB_HINT_LIGHT_PCT=$7 timeout 100 python tools/timeline_probe.py /tmp/x.json $1 $2 2>&1 | grep "^primary\|^shaded" | python -c "
import sys,json
for l in sys.stdin:
    k,_,j=l.partition(' '); d=json.loads(j)
    print('  ',k.ljust(15),'n1 %.4f  n8 max %.4f mean %.4f  noflush %.4f  x%.2f'%(d['n1']['max_ms'],d['n8']['max_ms'],d['n8']['mean_ms'],d['n8_no_flush_ms'],d['speedup_n8']), d['hint_stats_n8_part3'])
"
done; done
