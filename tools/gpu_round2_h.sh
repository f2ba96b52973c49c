mkdir -p gpurun_out
echo "== scheduling sweep (1 M soup, 16 spp)"
for t in 0 1 2 3 4 5 6; do PB2_TUNE=$t python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed "s/^/tune$t /"; done
for m in 12 16; do PB2_LIGHT_MINB=$m python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed "s/^/lightminb$m /"; done
