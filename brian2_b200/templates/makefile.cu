# b200 project: host code objects with $(CXX), everything touching the device with nvcc (sm_100a)
LIBRARY = {{library_name}}

SRCS = {{source_files}}
CU_SRCS = {{cu_source_files}}
H_SRCS = {{header_files}}
OBJS = ${SRCS:.cpp=.o}
OBJS := ${OBJS:.c=.o}
CU_OBJS = ${CU_SRCS:.cu=.o}
NVCC = {{nvcc}}
CUDA_HOME = {{cuda_home}}
OPTIMISATIONS = {{ compiler_flags }}
CXXFLAGS = -c -fPIC -fno-gnu-unique -Wno-write-strings $(OPTIMISATIONS) -I. -I$(CUDA_HOME)/include {{ compiler_debug_flags }}
NVCCFLAGS = -c {{ nvcc_flags }} -I. -Xcompiler -fPIC,-fno-gnu-unique {{ compiler_debug_flags }}
LFLAGS = -shared -fPIC {{ linker_flags }} {{ linker_debug_flags }} -L$(CUDA_HOME)/lib64 -lcudart_static -lpthread -ldl -lrt

all: $(LIBRARY)

.PHONY: all clean

$(LIBRARY): $(OBJS) $(CU_OBJS) makefile
	$(CXX) $(OBJS) $(CU_OBJS) -o $(LIBRARY) $(LFLAGS)

clean:
	{{ rm_cmd }}

%.o : %.cpp makefile $(H_SRCS)
	$(CXX) $(CXXFLAGS) $< -o $@

b200_kernels.o : b200_kernels.cu makefile $(H_SRCS) $(wildcard code_objects/*.cuh)
	$(NVCC) $(NVCCFLAGS) $< -o $@

%.o : %.cu makefile $(H_SRCS)
	$(NVCC) $(NVCCFLAGS) $< -o $@
