for b in 1 2 8; do echo "== B=$b"; timeout 100 python tools/time_layers.py $b 2>&1 | sed -n 2,3p | sed -e 's/stencil[^|]*|//'; done
