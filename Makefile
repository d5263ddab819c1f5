# mlsl-b200 build: one shared library (host runtime + sm_100a kernels), the launcher, C/C++ tests and examples.
#   make            -> mlsl_b200/lib/libmlsl_b200.so, bin/mlslrun, bin/* tests
#   make NO_CUDA=1  -> host-only library (no nvcc needed)
#   make sass       -> profiles/sass/*.sass listings of every kernel
CXX      ?= g++
NVCC     ?= /usr/local/cuda/bin/nvcc
CUDA_HOME ?= /usr/local/cuda
ARCH     := -gencode arch=compute_100a,code=sm_100a
CXXFLAGS := -O2 -g -std=c++17 -fPIC -Wall -Wextra -Wno-unused-parameter -pthread -Iinclude -Icsrc
NVFLAGS  := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-parameter -Iinclude -Icsrc \
            --expt-relaxed-constexpr -Xptxas -v
LDFLAGS  := -shared -pthread -lrt -ldl
ifdef DEBUG
CXXFLAGS += -O0 -DMLSLB_DEBUG
endif
ifdef TSAN
CXXFLAGS += -fsanitize=thread
LDFLAGS  += -fsanitize=thread
NO_CUDA := 1
endif

BUILD := build
LIBDIR := mlsl_b200/lib
LIB := $(LIBDIR)/libmlsl_b200.so

CORE_SRC := $(wildcard csrc/core/*.cpp)
CORE_OBJ := $(patsubst csrc/%.cpp,$(BUILD)/%.o,$(CORE_SRC))
ifdef NO_CUDA
CUDA_OBJ := $(BUILD)/cuda/backend_stub.o
else
CUDA_SRC := $(wildcard csrc/cuda/*.cu)
CUDA_OBJ := $(patsubst csrc/%.cu,$(BUILD)/%.o,$(CUDA_SRC))
LDFLAGS  += -L$(CUDA_HOME)/lib64 -lcudart_static
endif

# the host reduction loops are written to be auto-vectorised
$(BUILD)/core/host_backend.o: CXXFLAGS += -O3
$(BUILD)/core/net_backend.o: CXXFLAGS += -O3

TOOLS := bin/mlslrun
ifndef NO_CUDA
CUDA_EXAMPLES := bin/mlsl_example_cuda
endif
TESTS := $(CUDA_EXAMPLES) bin/libmlsl_quant_sample.so bin/quant_codec_check bin/mlsl_functional_test bin/cmlsl_smoke_test bin/cmlsl_functional_test bin/cmlsl_eplib_test bin/mlsl_sample bin/mlsl_example bin/mlsl_allreduce_bench

all: $(LIB) $(TOOLS) $(TESTS)

$(BUILD)/%.o: csrc/%.cpp $(wildcard csrc/core/*.hpp) include/mlsl.hpp include/mlsl.h include/eplib.h
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(BUILD)/%.o: csrc/%.cu $(wildcard csrc/core/*.hpp) $(wildcard csrc/cuda/*.cuh) $(wildcard csrc/cuda/*.hpp)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$(notdir $@).ptxas.log || (cat $(BUILD)/$(notdir $@).ptxas.log; false)

$(LIB): $(CORE_OBJ) $(CUDA_OBJ)
	@mkdir -p $(LIBDIR)
	$(CXX) -o $@ $^ $(LDFLAGS)

bin/mlslrun: csrc/tools/mlslrun.cpp
	@mkdir -p bin
	$(CXX) $(CXXFLAGS) -o $@ $<

bin/%: csrc/tests/%.cpp $(LIB)
	@mkdir -p bin
	$(CXX) $(CXXFLAGS) -o $@ $< -L$(LIBDIR) -lmlsl_b200 -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)'

bin/libmlsl_quant_sample.so: csrc/tests/quant_plugin_sample.c
	@mkdir -p bin
	gcc -O2 -g -std=gnu99 -Wall -shared -fPIC -o $@ $< -lm

bin/mlsl_example_cuda: csrc/tests/mlsl_example_cuda.cu $(LIB)
	@mkdir -p bin
	$(NVCC) -O2 -std=c++17 $(ARCH) -Iinclude -o $@ $< -L$(LIBDIR) -lmlsl_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../$(LIBDIR)'

bin/cmlsl_%: csrc/tests/cmlsl_%.c $(LIB)
	@mkdir -p bin
	gcc -O2 -g -std=gnu99 -Wall -Iinclude -o $@ $< -L$(LIBDIR) -lmlsl_b200 -lm -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)'

sass: $(LIB)
	@mkdir -p profiles/sass
	cuobjdump -sass $(LIB) > profiles/sass/libmlsl_b200.sass

# ThreadSanitizer run of the host runtime (ring, progress threads, shm protocol): 4 in-process ranks, hybrid 2x2
tsan: bin/mlslrun
	$(MAKE) CXX=/usr/bin/g++ TSAN=1 BUILD=/tmp/mlsl_tsan/build LIBDIR=/tmp/mlsl_tsan/lib LIB=/tmp/mlsl_tsan/lib/libmlsl_b200.so /tmp/mlsl_tsan/lib/libmlsl_b200.so
	/usr/bin/g++ -O1 -g -std=c++17 -fsanitize=thread -pthread -Iinclude -Icsrc csrc/tests/mlsl_functional_test.cpp -o /tmp/mlsl_tsan/ftest -L/tmp/mlsl_tsan/lib -lmlsl_b200 -Wl,-rpath,/tmp/mlsl_tsan/lib
	cd /tmp/mlsl_tsan && TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0" ./ftest 2 1 0 1 --inproc 4 | tail -1
	cd /tmp/mlsl_tsan && MLSL_NUM_SERVERS=2 MLSL_MSG_PRIORITY=1 TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0" ./ftest 1 0 1 0 --inproc 4 | tail -1
	# net backend (control server, receiver thread, TCP mesh): two launchers playing two nodes
	cd /tmp/mlsl_tsan && (TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0" $(CURDIR)/bin/mlslrun -n 2 --nnodes 2 --node-rank 1 --master-addr 127.0.0.1 --master-port 29877 ./ftest 2 1 > node1.out 2>&1 &) ; \
	  TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0" $(CURDIR)/bin/mlslrun -n 2 --nnodes 2 --node-rank 0 --master-addr 127.0.0.1 --master-port 29877 ./ftest 2 1 | grep -c "0 FAILED"
	# ... and with every two-level collective on its piece-by-piece route (4 KiB pieces)
	cd /tmp/mlsl_tsan && export MLSL_NET_HIER_KB=0 MLSL_NET_CHUNK_KB=4 TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0" && \
	  ($(CURDIR)/bin/mlslrun -n 2 --nnodes 2 --node-rank 1 --master-addr 127.0.0.1 --master-port 29883 ./ftest 2 1 > node1p.out 2>&1 &) ; \
	  $(CURDIR)/bin/mlslrun -n 2 --nnodes 2 --node-rank 0 --master-addr 127.0.0.1 --master-port 29883 ./ftest 2 1 | grep -c "0 FAILED" ; \
	  ! grep -q "ThreadSanitizer" node1p.out

# make install PREFIX=/opt/mlsl_b200: the layout of the reference's package (intel64/{bin,lib,include}, doc, examples,
# the environment script), plus the Python package
PREFIX ?= $(CURDIR)/_install
install: all
	@mkdir -p $(PREFIX)/intel64/bin $(PREFIX)/intel64/lib $(PREFIX)/intel64/include/mlsl $(PREFIX)/doc $(PREFIX)/examples $(PREFIX)/python
	cp $(LIB) $(PREFIX)/intel64/lib/
	ln -sf libmlsl_b200.so $(PREFIX)/intel64/lib/libmlsl.so   # drop-in name: -lmlsl and the reference's Python binding find it
	cp bin/mlslrun bin/libmlsl_quant_sample.so $(PREFIX)/intel64/bin/
	cp include/mlsl.hpp include/mlsl.h include/eplib.h $(PREFIX)/intel64/include/
	cp -r mlsl_b200 $(PREFIX)/python/ && rm -rf $(PREFIX)/python/mlsl_b200/__pycache__ $(PREFIX)/python/mlsl_b200/*/__pycache__
	cp README.md DESIGN.md docs/*.md $(PREFIX)/doc/
	cp examples/*.py csrc/tests/mlsl_example.cpp csrc/tests/mlsl_sample.cpp csrc/tests/mlsl_functional_test.cpp \
	   csrc/tests/cmlsl_smoke_test.c csrc/tests/quant_plugin_sample.c $(PREFIX)/examples/
	cp scripts/mlslvars.sh $(PREFIX)/intel64/bin/mlslvars.sh
	@echo "installed into $(PREFIX); source $(PREFIX)/intel64/bin/mlslvars.sh"

# the reference's `make testing`: the functional matrix on 4 local ranks (C++, C, Python, quantisation, net backend)
testing: all
	bash scripts/run_matrix.sh

# AddressSanitizer + UBSan run of the host runtime: in-process and multi-process functional test, user plug-in path
asan: bin/mlslrun bin/libmlsl_quant_sample.so
	$(MAKE) CXX=/usr/bin/g++ NO_CUDA=1 BUILD=/tmp/mlsl_asan/build LIBDIR=/tmp/mlsl_asan/lib LIB=/tmp/mlsl_asan/lib/libmlsl_b200.so \
	  CXXFLAGS="-O1 -g -std=c++17 -fPIC -pthread -Iinclude -Icsrc -fsanitize=address,undefined -fno-omit-frame-pointer" \
	  LDFLAGS="-shared -pthread -lrt -ldl -fsanitize=address,undefined" /tmp/mlsl_asan/lib/libmlsl_b200.so
	/usr/bin/g++ -O1 -g -std=c++17 -fsanitize=address,undefined -pthread -Iinclude -Icsrc csrc/tests/mlsl_functional_test.cpp -o /tmp/mlsl_asan/ftest -L/tmp/mlsl_asan/lib -lmlsl_b200 -Wl,-rpath,/tmp/mlsl_asan/lib
	cd /tmp/mlsl_asan && MLSL_BACKEND=host ./ftest 2 1 0 1 --inproc 4 | tail -1
	cd /tmp/mlsl_asan && MLSL_BACKEND=host MLSL_TEST_QUANT_LIB=$(CURDIR)/bin/libmlsl_quant_sample.so ./ftest 1 0 0 0 1 --inproc 4 | tail -1
	cd /tmp/mlsl_asan && MLSL_BACKEND=host MLSL_HEAP_SIZE_GB=0.2 $(CURDIR)/bin/mlslrun -n 4 ./ftest 2 1 | grep -c "0 FAILED"
	# net backend (tree broadcast, dissemination barrier, read-ahead reads, shared-memory rings): two launchers playing two nodes
	cd /tmp/mlsl_asan && ($(CURDIR)/bin/mlslrun -n 2 --nnodes 2 --node-rank 1 --master-addr 127.0.0.1 --master-port 29879 ./ftest 2 1 > node1.out 2>&1 &) ; \
	  $(CURDIR)/bin/mlslrun -n 2 --nnodes 2 --node-rank 0 --master-addr 127.0.0.1 --master-port 29879 ./ftest 2 1 | grep -c "0 FAILED" ; \
	  ! grep -q "AddressSanitizer\|runtime error" node1.out
	# ... and with every two-level collective on its piece-by-piece route (4 KiB pieces)
	cd /tmp/mlsl_asan && export MLSL_NET_HIER_KB=0 MLSL_NET_CHUNK_KB=4 && ($(CURDIR)/bin/mlslrun -n 2 --nnodes 2 --node-rank 1 --master-addr 127.0.0.1 --master-port 29881 ./ftest 2 1 > node1p.out 2>&1 &) ; \
	  $(CURDIR)/bin/mlslrun -n 2 --nnodes 2 --node-rank 0 --master-addr 127.0.0.1 --master-port 29881 ./ftest 2 1 | grep -c "0 FAILED" ; \
	  ! grep -q "AddressSanitizer\|runtime error" node1p.out

clean:
	rm -rf $(BUILD) $(LIB) bin _install

.PHONY: all clean sass tsan asan install testing
