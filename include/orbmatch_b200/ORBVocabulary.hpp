// ORBVocabulary.hpp -- C++ drop-in for ORB_SLAM3::ORBVocabulary::transform on the GPU.
//
// The reference's vocabulary type is
//     typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;   (include/ORBVocabulary.h)
// and the one member the matching hot path uses is
//     virtual void transform(const std::vector<TDescriptor> &features, BowVector &v, FeatureVector &fv, int levelsup) const;
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:146-147, :1127-1194), called with levelsup = 4 from Frame::ComputeBoW
// (Frame.cc:1005-1008) and KeyFrame::ComputeBoW (KeyFrame.cc:113).
//
// orbgpu::ORBVocabularyT DERIVES from the reference's class, so loading, saving, create(), score() and every other member stay the
// reference's own code; only this transform overload is overridden: the tree descent of every feature (FORB::distance, first child
// wins ties, node id at level L - levelsup), the BowVector accumulation (idf added once per occurrence in feature order, L1 norm
// summed in ascending word order, bit-exact doubles) and the FeatureVector run on the device (orbgpu_transform_descriptors), and the
// two std::maps are filled from the flat result.  There is no CPU fallback: weighting / scoring combinations the kernels do not
// implement throw.
//
//     #include "Thirdparty/DBoW2/DBoW2/FORB.h"
//     #include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"
//     #include <orbmatch_b200/ORBVocabulary.hpp>
//     typedef orbgpu::ORBVocabularyT<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;
#pragma once

#include <cstring>
#include <mutex>
#include <stdexcept>
#include <vector>

#include <orbmatch_b200.h>

#include <orbmatch_b200/ORBmatcher.hpp> // orbgpu::check, orbgpu::thread_context

namespace orbgpu
{
    template <class TDescriptor, class F>
    class ORBVocabularyT : public DBoW2::TemplatedVocabulary<TDescriptor, F>
    {
        typedef DBoW2::TemplatedVocabulary<TDescriptor, F> Base;

    public:
        // the reference's constructors (TemplatedVocabulary.h:64-87)
        ORBVocabularyT(int k = 10, int L = 5, DBoW2::WeightingType weighting = DBoW2::TF_IDF, DBoW2::ScoringType scoring = DBoW2::L1_NORM)
            : Base(k, L, weighting, scoring) {}
        virtual ~ORBVocabularyT()
        {
            if (dev_) orbgpu_voc_destroy(dev_);
        }

        using Base::transform; // the single-feature and BowVector-only overloads stay the reference's

        // TemplatedVocabulary.h:146-147 (definition :1127-1194)
        virtual void transform(const std::vector<TDescriptor> &features, DBoW2::BowVector &v, DBoW2::FeatureVector &fv, int levelsup) const
        {
            v.clear();
            fv.clear();
            if (this->empty()) return; // :1134-1137
            if (!((this->m_weighting == DBoW2::TF || this->m_weighting == DBoW2::TF_IDF) && this->m_scoring == DBoW2::L1_NORM))
                throw std::runtime_error("orbmatch_b200: transform on the GPU implements TF / TF_IDF weighting with L1 scoring (ORB-SLAM3's "
                                         "configuration, TemplatedVocabulary.h:57-58); there is no CPU fallback");
            orbgpu_ctx *ctx = thread_context();
            const orbgpu_voc *dv = device_copy(ctx);
            const int n = (int)features.size();
            if (n == 0) return;
            std::vector<uint8_t> desc((size_t)n * 32);
            for (int i = 0; i < n; i++) std::memcpy(&desc[(size_t)i * 32], features[i].template ptr<uint8_t>(), 32);
            std::vector<uint32_t> words(n), node_ids(n), feats(n);
            std::vector<double> values(n);
            std::vector<int32_t> offsets(n + 1);
            int32_t n_words = 0, n_nodes = 0;
            check(orbgpu_transform_descriptors(ctx, dv, n, desc.data(), levelsup, &n_words, words.data(), values.data(), &n_nodes, node_ids.data(),
                                               offsets.data(), feats.data()));
            // both results arrive in ascending key order: hinted inserts at the end are O(1)
            for (int j = 0; j < n_words; j++) v.insert(v.end(), std::make_pair((DBoW2::WordId)words[j], (DBoW2::WordValue)values[j]));
            for (int a = 0; a < n_nodes; a++)
            {
                DBoW2::FeatureVector::iterator it = fv.insert(fv.end(), std::make_pair((DBoW2::NodeId)node_ids[a], std::vector<unsigned int>()));
                it->second.assign(feats.begin() + offsets[a], feats.begin() + offsets[a + 1]);
            }
        }

    private:
        // flattens m_nodes (TemplatedVocabulary.h:303-335, :430) and uploads it once; a vocabulary that was (re)built or (re)loaded
        // since -- its node count changed -- is uploaded again
        const orbgpu_voc *device_copy(orbgpu_ctx *ctx) const
        {
            std::lock_guard<std::mutex> lock(mu_);
            const size_t nn = this->m_nodes.size();
            if (dev_ && dev_nodes_ == nn) return dev_;
            if (dev_) { orbgpu_voc_destroy(dev_); dev_ = nullptr; }
            std::vector<uint8_t> nd(nn * 32, 0);
            std::vector<int32_t> coff(nn + 1, 0);
            std::vector<uint32_t> cids, wid(nn, 0);
            std::vector<double> wt(nn, 0.0);
            for (size_t i = 0; i < nn; i++)
            {
                const typename Base::Node &node = this->m_nodes[i];
                if (!node.descriptor.empty()) std::memcpy(&nd[i * 32], node.descriptor.template ptr<uint8_t>(), 32);
                coff[i] = (int32_t)cids.size();
                for (size_t c = 0; c < node.children.size(); c++) cids.push_back((uint32_t)node.children[c]);
                wt[i] = node.weight;
                wid[i] = (uint32_t)node.word_id;
            }
            coff[nn] = (int32_t)cids.size();
            if (cids.empty()) cids.push_back(0);
            orbgpu_voc_host h;
            std::memset(&h, 0, sizeof(h));
            h.k = this->m_k; h.L = this->m_L; h.n_nodes = (int32_t)nn;
            h.node_desc = nd.data(); h.child_offsets = coff.data(); h.child_ids = cids.data(); h.weight = wt.data(); h.word_id = wid.data();
            check(orbgpu_voc_upload(ctx, &h, &dev_));
            dev_nodes_ = nn;
            return dev_;
        }
        mutable std::mutex mu_;
        mutable orbgpu_voc *dev_ = nullptr;
        mutable size_t dev_nodes_ = 0;
    };
} // namespace orbgpu
