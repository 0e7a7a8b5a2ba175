// rt/xorwow.hpp -- host restatement of cuRAND's XORWOW generator.
//
// The reference builds its scenes on the GPU from one cuRAND stream,
// curand_init(1984, 0, 0) (reference kernel.cu:105,184), drawing with
// curand_uniform (kernel.cu:157).  The render path here builds scenes on the
// host, so to get the *same* spheres the host needs the same numbers.
//
// cuRAND is not part of the reference tree; it ships with the CUDA Toolkit
// 12.9 this repo builds against.  Restated from its published header:
//   state init   curand_kernel.h:800-825  (_curand_init_inplace)
//   step         curand_kernel.h:863-874  (curand(curandStateXORWOW_t*))
//   to float     curand_uniform.h:69-72   (_curand_uniform)
// Subsequence 0 / offset 0 need no skip-ahead, which is all the scene stream
// uses.  The per-pixel render streams of the reference (subsequence =
// pixelIndex, a 2^67-stride skip) are NOT reproduced: the render path has its
// own counter-based stream (include/rt_rng.h).
#pragma once

#include <cstdint>

namespace rt {

class Xorwow {
public:
    explicit Xorwow(unsigned long long seed = 1984ULL)
    {
        const uint32_t s0 = static_cast<uint32_t>(seed) ^ 0xaad26b49u;
        const uint32_t s1 = static_cast<uint32_t>(seed >> 32) ^ 0xf7dcefddu;
        const uint32_t t0 = 1099087573u * s0;
        const uint32_t t1 = 2591861531u * s1;
        d_ = 6615241u + t1 + t0;
        v_[0] = 123456789u + t0;
        v_[1] = 362436069u ^ t0;
        v_[2] = 521288629u + t1;
        v_[3] = 88675123u ^ t1;
        v_[4] = 5783321u + t0;
    }

    uint32_t NextBits()
    {
        const uint32_t t = v_[0] ^ (v_[0] >> 2);
        v_[0] = v_[1];
        v_[1] = v_[2];
        v_[2] = v_[3];
        v_[3] = v_[4];
        v_[4] = (v_[4] ^ (v_[4] << 4)) ^ (t ^ (t << 1));
        d_ += 362437u;
        ++draws_;
        return v_[4] + d_;
    }

    // (0,1], fp32, exactly curand_uniform.
    float Uniform() { return static_cast<float>(NextBits()) * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }

    unsigned long long Count() const { return draws_; }

private:
    uint32_t v_[5];
    uint32_t d_;
    unsigned long long draws_ = 0;
};

} // namespace rt
