for b in 1 2 4 8 16; do echo "== B=$b"; timeout 100 python tools/time_layers.py $b 2>&1 | sed -n 2,3p | sed -e 's/stencil[^|]*|//'; done
echo "== B=2 skip all"; L3D_C3_DEBUG_SKIP=7 timeout 100 python tools/time_layers.py 2 2>&1 | sed -n 2,3p | sed -e 's/stencil[^|]*|//'
