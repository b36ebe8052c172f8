# builds variants/lib_timers.so: the product library with -DRC_TIMERS (phase timers, dev tool)
mkdir -p variants
cd gerrychain_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DRC_TIMERS -o ../../variants/lib_timers.so recom_abi.cu
