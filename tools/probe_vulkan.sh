#!/bin/bash
# Probe a GPU box for anything that could run the reference's Vulkan path: loader, ICDs, tools.
# Output is committed under profiles/ so the "reference-Vulkan comparator not measurable here" statement is checkable.
echo "== date: $(date -u +%FT%TZ)"
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,driver_version --format=csv,noheader 2>&1 | head -8
echo "== ICD manifests"; ls -la /usr/share/vulkan/icd.d /etc/vulkan/icd.d /usr/share/vulkan/implicit_layer.d /usr/local/share/vulkan/icd.d 2>&1
echo "== NVIDIA ICD json anywhere"; find / -xdev \( -name "nvidia_icd*.json" -o -name "lvp_icd*.json" -o -name "*_icd.*.json" \) 2>/dev/null | head
echo "== loader / driver libs"; ldconfig -p | grep -i -E "vulkan|libGLX_nvidia|libnvidia-glcore|libnvidia-gpucomp|libEGL_nvidia|libnvidia-vulkan" 
find / -xdev \( -name "libvulkan.so*" -o -name "libGLX_nvidia.so*" -o -name "libvulkan_lvp.so*" -o -name "libnvidia-vulkan-producer.so*" \) 2>/dev/null | head
echo "== tools"; for t in vulkaninfo glslangValidator glslc spirv-as Xvfb cmake; do printf "%s: " $t; which $t || echo "absent"; done
echo "== headers"; ls /usr/include/vulkan 2>&1 | head -3; ls /usr/include/glm 2>&1 | head -2; ls /usr/include/GLFW 2>&1 | head -2
echo "== python vulkan bindings"; python -c "import vulkan" 2>&1 | tail -1
echo "== NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"
echo "== host"; nproc; free -g | head -2; nvidia-smi topo -m 2>&1 | head -12
