# Convenience targets for C users.  The authoritative build is `python __graft_entry__.py`
# (mini-nbody_b200/build.py); this Makefile runs the same commands (the loop re-scheduler is a Python step in both).
NVCC   ?= nvcc
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
PKG    := mini-nbody_b200
CSRC   := $(PKG)/csrc
OBJS   := $(PKG)/build/force_f32.o $(PKG)/build/force_f64.o $(PKG)/build/integrate.o $(PKG)/build/step_small.o $(PKG)/build/capi.o

all: $(PKG)/libnbody_b200.so apps/nbody oracle

$(PKG)/build/%.o: $(CSRC)/%.cu $(CSRC)/nbody_internal.cuh $(CSRC)/force_f32_inner.cuh $(CSRC)/stream.cuh include/nbody.h
	@mkdir -p $(PKG)/build
	$(NVCC) $(NVFLAGS) -c $< -o $@

# link, then the post-ptxas loop re-scheduler (same step as build.py: patch a copy, verify, replace; a loop that cannot be
# patched keeps ptxas's schedule and is reported).  NBODY_B200_NO_SCHED=1 skips the patching.
$(PKG)/libnbody_b200.so: $(OBJS) $(PKG)/sass_sched.py $(PKG)/sass_check.py
	$(NVCC) -shared -o $@ $(OBJS) -ldl
	python3 $(PKG)/build.py --resched-only

apps/nbody: apps/nbody.c include/nbody.h $(PKG)/libnbody_b200.so
	gcc -std=c11 -O2 -D_POSIX_C_SOURCE=200809L $< -o $@ -L$(PKG) -lnbody_b200 -Wl,-rpath,$(abspath $(PKG)) -Wl,-rpath,'$$ORIGIN/../$(PKG)'

oracle:
	$(MAKE) -C oracle

test:
	python -m pytest tests -q -m "not gpu"

clean:
	rm -rf $(PKG)/build $(PKG)/libnbody_b200.so apps/nbody oracle/_build

.PHONY: all oracle test clean
