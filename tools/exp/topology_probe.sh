nvidia-smi topo -m 2>&1 | head -14 | cut -c1-200
lscpu | grep -i "numa\|socket\|^CPU(s)\|model name" | head -8
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then echo "$(basename $d) numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist) width=$(cat $d/current_link_width) speed=$(cat $d/current_link_speed)"; fi; done
python - <<'PY'
import torch
for i in range(torch.cuda.device_count()):
    p=torch.cuda.get_device_properties(i)
    print(i, getattr(p,'pci_bus_id',None), getattr(p,'pci_device_id',None), getattr(p,'pci_domain_id',None))
PY
free -g | head -2
