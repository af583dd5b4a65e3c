for wf in 0 1 2 3 4; do echo "== wavefront $wf"; timeout 200 python tools/fused_probe.py c2 4 g,d wavefront=$wf 2>&1 | grep "^c2\|rel" | cut -c1-330; done
for wf in 0 2 3 4; do echo "== wavefront $wf"; timeout 300 python tools/fused_probe.py c3s 2 g,d wavefront=$wf 2>&1 | grep "^c3s\|rel" | cut -c1-330; done
