// stand-in for TBB: the reference only uses tbb::parallel_for outside MAKE_DETERMINISTIC builds; serial loop.
#pragma once
namespace tbb {
template <class Index, class F>
void parallel_for(Index first, Index last, const F& f)
{
    for (Index i = first; i < last; ++i) f(i);
}
}  // namespace tbb
