mkdir -p gpurun_out
echo "== probes: digit loop vs digit tables"
python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16"
PB2_HALTON_TAB=1 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed 's/^/tables /'
python tools/probe.py file tests/scenes/killeroo_like.pbrt 16 "0" 2>&1 | grep "probe file"
PB2_HALTON_TAB=1 python tools/probe.py file tests/scenes/killeroo_like.pbrt 16 "0" 2>&1 | grep "probe file" | sed 's/^/tables /'
echo "== gpu tests with the tables"; PB2_HALTON_TAB=1 timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
