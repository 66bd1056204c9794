// Replacement of moldyn/Clustering's src/density_clustering_cuda.hpp inside the reference's source tree.
//
// The reference's call sites include "density_clustering_cuda.hpp" by name (src/density_clustering.cpp:31-33,
// src/clustering.cpp:38-40); a quoted include looks in the including file's own directory first, so the binding is
// made by putting THIS file there (cmake/dcb200.cmake copies it over the original at configure time).  It pulls the
// reference's own types (Pops, Neighborhood, FreeEnergy) from the reference's headers, exactly as the original does
// (density_clustering_cuda.hpp:3-4), and then the Clustering::Density::CUDA functions from <dcb200/density_cuda.hpp>,
// which forward to libdcb200.so.  src/density_clustering_cuda.cu and src/density_clustering_cuda_kernels.cu are no
// longer compiled.
#pragma once

#include "config.hpp"
#include "density_clustering_common.hpp"

#include <dcb200/density_cuda.hpp>
