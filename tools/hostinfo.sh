#!/bin/bash
# Box facts for the host-side design: cores, NUMA, GPU topology.
echo "== nproc: $(nproc)"; lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|MHz' 
(numactl -H 2>/dev/null || echo "no numactl")
nvidia-smi topo -m 2>/dev/null | head -30
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done
free -g | head -2
