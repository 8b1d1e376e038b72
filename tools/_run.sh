for i in 1 2; do
for kb in 16 32 64; do
echo "min_kb=$kb $(GT_SPLITK_MIN_KB=$kb python tools/graph_trace.py molpcba 2>&1 | head -1 | cut -c1-60)"
done
done
for kb in 16 32 64; do
echo "code2 min_kb=$kb $(GT_SPLITK_MIN_KB=$kb python tools/graph_trace.py code2 2>&1 | head -1 | cut -c1-60)"
done
