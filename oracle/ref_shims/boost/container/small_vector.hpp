// Build shim (NOT reference code): the reference's fill_voxels_cpu.cc includes
// boost::container::small_vector, and boost is not installed here.  A
// std::vector with the same template signature is a drop-in for its use there.
#pragma once
#include <cstddef>
#include <vector>
namespace boost { namespace container {
template <class T, std::size_t N>
class small_vector : public std::vector<T> {
 public:
  using std::vector<T>::vector;
};
}}  // namespace boost::container
