# CUDA build wiring of moldyn/Clustering on libdcb200.so -- replaces CMakeLists.txt:110-137 of the reference
# (find_package(CUDA), the -arch=compute_30 nvcc flags, the two .cu sources and cuda_add_executable).
#
# In the reference's CMakeLists.txt:
#
#     if(${USE_CUDA})
#       set(DCB200_ROOT "/path/to/this/repository")
#       include(${DCB200_ROOT}/cmake/dcb200.cmake)          # instead of lines 110-127
#     endif()
#     ...
#     add_executable(${PROGNAME} ${CLUSTERING_SRCS})        # instead of lines 133-137: no cuda_add_executable,
#     target_link_libraries(${PROGNAME} ${CLUSTERING_LIBS}) # nvcc is not needed to build the reference any more
#
# What it does:
#   * builds (once) or locates libdcb200.so -- the library is compiled by its own Makefile with nvcc for sm_100a;
#   * puts the forwarding header in place of src/density_clustering_cuda.hpp (the call sites include it by name and a
#     quoted include searches the including file's directory first), keeping the original as *.reference;
#   * adds include/ to the include path, -DUSE_CUDA, and the library + an rpath to CLUSTERING_LIBS.
# The call sites (src/density_clustering.cpp:113-118, :616-621, :659-663, :716-720, :746-750, :808-814 and
# src/clustering.cpp:110-113) compile unchanged; tests/test_reference_callsites.py checks exactly that.

if(NOT DEFINED DCB200_ROOT)
  get_filename_component(DCB200_ROOT "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
endif()
message(STATUS "using CUDA through libdcb200 (${DCB200_ROOT})")

set(DCB200_LIBRARY "${DCB200_ROOT}/clustering_b200/libdcb200.so")
if(NOT EXISTS "${DCB200_LIBRARY}")
  message(STATUS "building libdcb200.so (nvcc, sm_100a)")
  execute_process(COMMAND make -C "${DCB200_ROOT}/clustering_b200/csrc" -j8 lib RESULT_VARIABLE DCB200_BUILD_RC)
  if(NOT DCB200_BUILD_RC EQUAL 0)
    message(FATAL_ERROR "could not build ${DCB200_LIBRARY}")
  endif()
endif()

set(DCB200_CUDA_HEADER "${CMAKE_CURRENT_SOURCE_DIR}/src/density_clustering_cuda.hpp")
if(EXISTS "${DCB200_CUDA_HEADER}" AND NOT EXISTS "${DCB200_CUDA_HEADER}.reference")
  file(RENAME "${DCB200_CUDA_HEADER}" "${DCB200_CUDA_HEADER}.reference")
endif()
configure_file("${DCB200_ROOT}/include/dcb200/reference_tree/density_clustering_cuda.hpp" "${DCB200_CUDA_HEADER}" COPYONLY)

include_directories("${DCB200_ROOT}/include")
add_definitions(-DUSE_CUDA)
set(CLUSTERING_LIBS ${CLUSTERING_LIBS} "${DCB200_LIBRARY}")
set(CMAKE_INSTALL_RPATH "${DCB200_ROOT}/clustering_b200")
set(CMAKE_BUILD_WITH_INSTALL_RPATH TRUE)
