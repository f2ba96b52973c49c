mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== envlight first"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "envlight" 2>&1 | tail -12
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
