// ORACLE HARNESS (test infrastructure, NOT product code).
//
// Drives the reference's OWN MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:444-535), compiled unmodified where it
// lies under /root/reference together with the reference's real include/MapPoint.h, against the stand-in classes of
// shim_mp/mp_stubs.h.  Pins oracle_compute_distinctive_descriptors (and through it the GPU kernel).
//
// The reference collects the observations from a std::map keyed by KeyFrame POINTER (:453, :467), so the order of
// vDescriptors -- and with it which of several equally good descriptors wins -- is the address order of the key frames.
// The harness allocates the key frames of one map point as one array, so address order == observation order of the caller.
#include "../include/MapPoint.h"

#include <cstdint>
#include <cstring>
#include <vector>

using namespace ORB_SLAM3;

// For map point p with observations [offsets[p], offsets[p+1]): out_desc[p] = the descriptor the reference keeps (32 bytes;
// untouched when the point has no usable observation), best_idx[p] = position of the FIRST observation that carries exactly
// that descriptor (the reference's BestIdx: the scan at :511-526 takes the first index among equal medians, and equal
// descriptors have equal medians), -1 when nothing was selected.  kf_bad (may be null): key frames to flag isBad() (:471).
extern "C" int ref_compute_distinctive(int32_t n_mp, const int32_t *offsets, const uint8_t *desc, const uint8_t *kf_bad,
                                       int32_t *best_idx, uint8_t *out_desc)
{
    for (int32_t p = 0; p < n_mp; p++) {
        const int32_t s = offsets[p], n = offsets[p + 1] - s;
        best_idx[p] = -1;
        std::vector<KeyFrame> kfs((size_t)(n > 0 ? n : 0));
        MapPoint mp;
        for (int32_t i = 0; i < n; i++) {
            KeyFrame &K = kfs[(size_t)i];
            K.mnId = (unsigned long)i;
            K.mDescriptors = cv::Mat(1, 32, CV_8U);
            std::memcpy(K.mDescriptors.data, desc + (size_t)(s + i) * 32, 32);
            K.mvuRight.assign(1, -1.f);
            K.bad = kf_bad && kf_bad[s + i];
            mp.AddObservation(&K, 0); // MapPoint.cc:168-205
        }
        mp.ComputeDistinctiveDescriptors();
        cv::Mat d = mp.GetDescriptor();
        if (d.empty()) continue;
        std::memcpy(out_desc + (size_t)p * 32, d.data, 32);
        for (int32_t i = 0; i < n; i++) {
            if (kfs[(size_t)i].bad) continue;
            if (std::memcmp(desc + (size_t)(s + i) * 32, d.data, 32) == 0) {
                best_idx[p] = i;
                break;
            }
        }
    }
    return 0;
}
