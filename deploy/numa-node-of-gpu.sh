#!/bin/sh
# Prints the NUMA node a GPU ordinal is attached to (sysfs of its PCI device); 0 when the platform does not say.
# Used by deploy/paladin-worker@.service.d/10-gpu.conf to bind worker i and its pinned staging buffers next to GPU i.
gpu="${1:-0}"
bus=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i "$gpu" 2>/dev/null | tr 'A-Z' 'a-z' | sed 's/^0000//')
node=$(cat "/sys/bus/pci/devices/${bus}/numa_node" 2>/dev/null || echo 0)
[ "$node" -lt 0 ] 2>/dev/null && node=0
echo "${node:-0}"
