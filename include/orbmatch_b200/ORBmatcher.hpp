// ORBmatcher.hpp -- C++ drop-in adapter: the reference's ORB_SLAM3::ORBmatcher call signatures
// (include/ORBmatcher.h:37-84) on top of the C ABI of include/orbmatch_b200.h.
//
// It only packs the members the reference matcher reads into the flat host structs, calls the
// GPU library, and scatters the results back into the caller's STL containers in the reference's
// layouts (vnMatches12, vpMapPointMatches, vMatchedPairs, F.mvpMapPoints).  No matching logic lives
// here and there is no CPU fallback: a failing GPU call throws std::runtime_error.
//
// The class is a template over the reference's own Frame / KeyFrame / MapPoint types so that it
// compiles inside ORB-SLAM3 unchanged:
//     #include <orbmatch_b200/ORBmatcher.hpp>
//     using ORBmatcherGPU = orbgpu::ORBmatcherT<ORB_SLAM3::Frame, ORB_SLAM3::KeyFrame, ORB_SLAM3::MapPoint>;
// and against the oracle's stub types for the parity tests (tests/cpp/adapter_parity.cc).
//
// Covered overloads: SearchForInitialization, SearchByProjection x5 (local map points;
// Cur/Last; Cur/KF relocalisation; KF/Sim3 with and without the keyframe list), SearchByBoW x2, SearchForTriangulation
// (monocular pinhole), Fuse x2, SearchBySim3, DescriptorDistance.  The overloads that project on their own evaluate
// the per-point prologue (pose transform, projection, frustum / distance / viewing gates, PredictScale) HERE, with
// the host's own Sophus / Eigen and in the reference's operation order; the window search, the Hamming distances, the
// ordered "keypoint already taken" rule, the acceptance threshold and the rotation histogram run on the GPU
// (orbgpu_search_projected).  The stereo-fisheye branches (Nleft != -1, bRight) are not on the GPU hot path.
#pragma once

#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <memory>
#include <vector>

#include <orbmatch_b200.h> // compile with -I<repo>/include

namespace orbgpu
{
    inline void check(int rc)
    {
        if (rc != ORBGPU_OK) throw std::runtime_error(std::string("orbmatch_b200: ") + orbgpu_last_error());
    }

    // one context (stream + workspace) per calling thread: the reference calls the matcher from the
    // Tracking, LocalMapping and LoopClosing threads concurrently (System.cc:234,254)
    inline orbgpu_ctx *thread_context(int device = 0)
    {
        struct Holder
        {
            orbgpu_ctx *ctx = nullptr;
            ~Holder() { if (ctx) orbgpu_destroy(ctx); }
        };
        thread_local Holder h;
        if (!h.ctx) check(orbgpu_create(device, &h.ctx));
        return h.ctx;
    }

    // flat copy of the members of a Frame / KeyFrame that the matcher reads
    struct PackedFrame
    {
        std::vector<uint8_t> desc;
        std::vector<float> xy, angle, u_right, sf, s2;
        std::vector<int32_t> octave;
        std::vector<uint32_t> fv_nodes, fv_feats;
        std::vector<int32_t> fv_off;
        orbgpu_frame_host h;

        template <class FS>
        void pack(const FS &F, float minX, float minY, float maxX, float maxY, bool with_uright)
        {
            const int n = F.N;
            desc.resize((size_t)n * 32);
            xy.resize((size_t)n * 2);
            angle.resize(n);
            octave.resize(n);
            for (int i = 0; i < n; i++)
            {
                std::memcpy(&desc[(size_t)i * 32], F.mDescriptors.row(i).template ptr<uint8_t>(), 32);
                xy[2 * i] = F.mvKeysUn[i].pt.x;
                xy[2 * i + 1] = F.mvKeysUn[i].pt.y;
                octave[i] = F.mvKeysUn[i].octave;
                angle[i] = F.mvKeysUn[i].angle;
            }
            u_right.clear();
            if (with_uright) u_right.assign(F.mvuRight.begin(), F.mvuRight.end());
            sf.assign(F.mvScaleFactors.begin(), F.mvScaleFactors.end());
            s2.assign(F.mvLevelSigma2.begin(), F.mvLevelSigma2.end());
            fv_nodes.clear(); fv_feats.clear(); fv_off.clear();
            for (const auto &kv : F.mFeatVec) // std::map<NodeId, vector<uint>>: ascending node id, ascending feature id
            {
                fv_nodes.push_back(kv.first);
                fv_off.push_back((int32_t)fv_feats.size());
                fv_feats.insert(fv_feats.end(), kv.second.begin(), kv.second.end());
            }
            fv_off.push_back((int32_t)fv_feats.size());
            std::memset(&h, 0, sizeof(h));
            h.n = n;
            h.desc = desc.data(); h.kp_xy = xy.data(); h.octave = octave.data(); h.angle = angle.data();
            h.u_right = u_right.empty() ? nullptr : u_right.data();
            h.min_x = minX; h.min_y = minY; h.max_x = maxX; h.max_y = maxY;
            h.grid_inv_w = F.mfGridElementWidthInv; h.grid_inv_h = F.mfGridElementHeightInv;
            h.grid_cols = ORBGPU_FRAME_GRID_COLS; h.grid_rows = ORBGPU_FRAME_GRID_ROWS;
            h.n_levels = (int32_t)sf.size();
            h.scale_factors = sf.data(); h.level_sigma2 = s2.data();
            h.fv_n_nodes = (int32_t)fv_nodes.size();
            h.fv_node_ids = fv_nodes.data(); h.fv_offsets = fv_off.data(); h.fv_features = fv_feats.data();
        }
    };

    struct DeviceFrameGuard
    {
        orbgpu_frame *f = nullptr;
        DeviceFrameGuard(orbgpu_ctx *ctx, const orbgpu_frame_host &h) { check(orbgpu_frame_upload(ctx, &h, &f)); }
        ~DeviceFrameGuard() { orbgpu_frame_destroy(f); }
        DeviceFrameGuard(const DeviceFrameGuard &) = delete;
        DeviceFrameGuard &operator=(const DeviceFrameGuard &) = delete;
    };

    // Device frames are cached per calling thread, keyed on the reference's own identity of the object: Frame::mnId / KeyFrame::mnId
    // (Frame.h:278, KeyFrame.h:243; key points and descriptors are immutable after construction, KeyFrame.h:378-400) plus what can
    // still change afterwards -- the FeatureVector (ComputeBoW runs once, later) and whether mvuRight is needed.  A repeated call on
    // the same Frame / KeyFrame (TrackReferenceKeyFrame then TrackLocalMap on the current frame; a key frame matched against each of
    // its neighbours in CreateNewMapPoints / SearchInNeighbors / loop detection) skips the pack + upload + grid build, which cost
    // more than the search itself.  Bounded LRU; set_frame_cache_capacity(0) turns it off.  Entries referenced by a call in progress
    // (a batched search holds K + 1 of them) are pinned: eviction skips them, and the cache may exceed its capacity until they are released.
    struct FrameCache
    {
        struct Key
        {
            int kind; // 0 Frame, 1 KeyFrame
            unsigned long id;
            int n, with_uright;
            size_t fv_nodes;
            bool operator==(const Key &o) const { return kind == o.kind && id == o.id && n == o.n && with_uright == o.with_uright && fv_nodes == o.fv_nodes; }
        };
        struct Entry { Key key; orbgpu_frame *f; unsigned long stamp; int pins; };
        std::vector<Entry> entries;
        unsigned long clock = 0, hits = 0, misses = 0;
        size_t capacity = 48;
        ~FrameCache() { clear(); }
        void clear()
        {
            for (Entry &e : entries) orbgpu_frame_destroy(e.f);
            entries.clear();
        }
        orbgpu_frame *find(const Key &k)
        {
            for (Entry &e : entries)
                if (e.key == k) { e.stamp = ++clock; e.pins++; hits++; return e.f; }
            misses++;
            return nullptr;
        }
        void unpin(const orbgpu_frame *f)
        {
            for (Entry &e : entries)
                if (e.f == f) { if (e.pins > 0) e.pins--; return; }
        }
        void insert(const Key &k, orbgpu_frame *f)
        {
            // an older copy of the same object (its FeatureVector has been computed since) is dropped, then the least recently used
            for (size_t i = 0; i < entries.size();)
                if (entries[i].key.kind == k.kind && entries[i].key.id == k.id && entries[i].pins == 0) { orbgpu_frame_destroy(entries[i].f); entries.erase(entries.begin() + i); }
                else i++;
            while (entries.size() >= capacity)
            {
                size_t lru = entries.size();
                for (size_t i = 0; i < entries.size(); i++)
                    if (entries[i].pins == 0 && (lru == entries.size() || entries[i].stamp < entries[lru].stamp)) lru = i;
                if (lru == entries.size()) break; // everything is in use by the call in progress
                orbgpu_frame_destroy(entries[lru].f);
                entries.erase(entries.begin() + lru);
            }
            entries.push_back(Entry{k, f, ++clock, 1});
        }
    };
    // the cache lives and dies with the thread's context (its frames were uploaded on that context's stream)
    inline FrameCache &frame_cache()
    {
        thread_local FrameCache c;
        return c;
    }
    inline void set_frame_cache_capacity(size_t n)
    {
        frame_cache().capacity = n;
        if (n == 0) frame_cache().clear();
    }
    // device copy of a Frame (kind 0) or KeyFrame (kind 1): from the thread's cache, or packed + uploaded now
    struct FrameRef
    {
        orbgpu_frame *f = nullptr;
        bool owned = false;
        template <class FS>
        FrameRef(orbgpu_ctx *ctx, int kind, const FS &F, float minX, float minY, float maxX, float maxY, bool with_uright)
        {
            FrameCache &c = frame_cache();
            const FrameCache::Key k{kind, (unsigned long)F.mnId, (int)F.N, with_uright ? 1 : 0, F.mFeatVec.size()};
            if (c.capacity > 0 && (f = c.find(k))) return;
            PackedFrame p;
            p.pack(F, minX, minY, maxX, maxY, with_uright);
            check(orbgpu_frame_upload(ctx, &p.h, &f));
            if (c.capacity > 0) c.insert(k, f);
            else owned = true;
        }
        ~FrameRef()
        {
            if (owned) orbgpu_frame_destroy(f);
            else if (f) frame_cache().unpin(f);
        }
        FrameRef(const FrameRef &) = delete;
        FrameRef &operator=(const FrameRef &) = delete;
    };
    // Checker of the device-side Frame::isInFrustum (north_star: "any window-boundary candidate disagreement is reported, with a target
    // of zero"): when enabled, ORBmatcherT::SearchLocalPoints also runs the reference's own Frame::isInFrustum on the host for every
    // point it tested on the device and counts the points whose result differs.
    struct FrustumReport
    {
        bool enabled = false;
        unsigned long points = 0, in_view_mismatch = 0, level_mismatch = 0, projection_mismatch = 0;
    };
    inline FrustumReport &frustum_report()
    {
        static FrustumReport r;
        return r;
    }


    // points already projected into the target frame (inputs of orbgpu_search_projected); index m == source index
    struct ProjBatch
    {
        std::vector<uint8_t> desc, active, locks;
        std::vector<float> uv, radius, ur, angle;
        std::vector<int32_t> lo, hi;
        explicit ProjBatch(int n) : desc((size_t)n * 32, 0), active(n, 0), locks(n, 1), uv((size_t)n * 2, 0.f), radius(n, 0.f), ur(n, 0.f),
                                    angle(n, 0.f), lo(n, -1), hi(n, -1) {}
        template <class Mat>
        void set(int m, const Mat &d, float u, float v, float r, int min_level, int max_level)
        {
            std::memcpy(&desc[(size_t)m * 32], d.template ptr<uint8_t>(), 32);
            uv[2 * m] = u; uv[2 * m + 1] = v; radius[m] = r; lo[m] = min_level; hi[m] = max_level; active[m] = 1;
        }
        orbgpu_projpoints_host host() const
        {
            orbgpu_projpoints_host h;
            std::memset(&h, 0, sizeof(h));
            h.n = (int32_t)active.size();
            h.desc = desc.data(); h.uv = uv.data(); h.radius = radius.data(); h.min_level = lo.data(); h.max_level = hi.data();
            h.ur = ur.data(); h.active = active.data(); h.locks = locks.data(); h.angle = angle.data();
            return h;
        }
    };
    struct ProjResult
    {
        std::vector<int32_t> best_idx, best_dist, kp_owner;
        int32_t nmatches = 0;
    };
    inline ProjResult run_projected(orbgpu_ctx *ctx, const orbgpu_frame *f, int n_kp, const ProjBatch &b, float max_dist, bool ordered,
                                    bool stereo_gate, bool chi2_gate, bool check_ori, const std::vector<float> *inv_sigma2,
                                    const std::vector<uint8_t> *kp_locked)
    {
        ProjResult r;
        const int M = (int)b.active.size();
        r.best_idx.assign(M > 0 ? M : 1, -1);
        r.best_dist.assign(M > 0 ? M : 1, 256);
        r.kp_owner.assign(n_kp > 0 ? n_kp : 1, -1);
        orbgpu_projsearch_params prm;
        std::memset(&prm, 0, sizeof(prm));
        prm.max_dist = max_dist; prm.ordered = ordered; prm.stereo_gate = stereo_gate; prm.chi2_gate = chi2_gate; prm.check_ori = check_ori;
        prm.inv_level_sigma2 = inv_sigma2 ? inv_sigma2->data() : nullptr;
        const orbgpu_projpoints_host h = b.host();
        check(orbgpu_search_projected(ctx, f, &h, &prm, kp_locked ? kp_locked->data() : nullptr, r.best_idx.data(), r.best_dist.data(),
                                      r.kp_owner.data(), &r.nmatches));
        return r;
    }

    template <class FrameT, class KeyFrameT, class MapPointT>
    class ORBmatcherT
    {
    public:
        // ORBmatcher.h:37
        ORBmatcherT(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

        // ORBmatcher.h:40 (ORBmatcher.cc:2388-2408)
        template <class Mat>
        static int DescriptorDistance(const Mat &a, const Mat &b)
        {
            int32_t d = 0;
            check(orbgpu_descriptor_distance(thread_context(), 1, a.template ptr<uint8_t>(), b.template ptr<uint8_t>(), &d));
            return d;
        }

        // SearchByNN: named in the matcher family but not present in this fork (no declaration in ORBmatcher.h) -- defined by this
        // repo as the brute-force form of SearchByBoW's inner loop (ORBmatcher.cc:327-355) with its accept rule (:392-395):
        // for every row of `queries` the nearest row of `database` (first index among equal distances) is accepted when
        // best <= th && (float)best < mfNNratio * (float)second.  vnMatches = vector<int>(queries.rows, -1) then filled;
        // returns the number of matches.  Both matrices are N x 32 CV_8U like Frame::mDescriptors.
        template <class Mat>
        int SearchByNN(const Mat &queries, const Mat &database, std::vector<int> &vnMatches, int th = 50 /* TH_LOW */)
        {
            orbgpu_ctx *ctx = thread_context();
            const int nq = queries.rows, nd = database.rows;
            vnMatches.assign((size_t)nq, -1);
            if (nq == 0 || nd == 0) return 0;
            // the device database lives with the calling thread and is re-used while it is large enough: a repeated call re-uploads the
            // rows in chunks on the database's own stream and searches every chunk as it arrives (orbgpu_knn2_ratio_update)
            struct DbHolder
            {
                orbgpu_db *db = nullptr;
                int64_t capacity = 0;
                ~DbHolder() { if (db) orbgpu_db_destroy(db); }
            };
            thread_local DbHolder holder;
            std::vector<int32_t> bi((size_t)nq), bd((size_t)nq), sd((size_t)nq), m((size_t)nq);
            if (holder.capacity < nd)
            {
                if (holder.db) orbgpu_db_destroy(holder.db);
                holder.db = nullptr;
                holder.capacity = 0;
                check(orbgpu_db_upload(ctx, nd, database.template ptr<uint8_t>(), &holder.db));
                holder.capacity = nd;
                check(orbgpu_knn2_ratio(ctx, holder.db, nq, queries.template ptr<uint8_t>(), th, mfNNratio, bi.data(), bd.data(), sd.data(), m.data()));
            }
            else
                check(orbgpu_knn2_ratio_update(ctx, holder.db, nd, database.template ptr<uint8_t>(), nq, queries.template ptr<uint8_t>(), th, mfNNratio,
                                               bi.data(), bd.data(), sd.data(), m.data()));
            int nmatches = 0;
            for (int i = 0; i < nq; i++) {
                vnMatches[(size_t)i] = m[(size_t)i];
                nmatches += m[(size_t)i] >= 0;
            }
            return nmatches;
        }

        // ORBmatcher.h:69 (ORBmatcher.cc:735-878)
        template <class Point2f>
        int SearchForInitialization(FrameT &F1, FrameT &F2, std::vector<Point2f> &vbPrevMatched, std::vector<int> &vnMatches12,
                                    int windowSize = 10)
        {
            orbgpu_ctx *ctx = thread_context();
            FrameRef d1(ctx, 0, F1, F1.mnMinX, F1.mnMinY, F1.mnMaxX, F1.mnMaxY, false), d2(ctx, 0, F2, F2.mnMinX, F2.mnMinY, F2.mnMaxX, F2.mnMaxY, false);
            const int n1 = F1.N;
            std::vector<float> prev((size_t)n1 * 2);
            for (int i = 0; i < n1; i++) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
            std::vector<int32_t> m12(n1 > 0 ? n1 : 1);
            int32_t nmatches = 0;
            check(orbgpu_search_for_initialization(ctx, d1.f, d2.f, prev.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0,
                                                   m12.data(), &nmatches));
            vnMatches12.assign(m12.begin(), m12.begin() + n1); // vnMatches12 = vector<int>(N1, -1) then filled (:739)
            for (int i = 0; i < n1; i++) { vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
            return nmatches;
        }

        // ORBmatcher.h:44 (ORBmatcher.cc:44-242), monocular / RGB-D path (F.Nleft == -1)
        int SearchByProjection(FrameT &F, const std::vector<MapPointT *> &vpMapPoints, const float th = 3, const bool bFarPoints = false,
                               const float thFarPoints = 50.0f)
        {
            if (F.Nleft != -1) throw std::runtime_error("orbmatch_b200: stereo-fisheye (Nleft != -1) path is not on the GPU hot path");
            orbgpu_ctx *ctx = thread_context();
            bool any_right = false;
            for (int i = 0; i < F.N && !any_right; i++) any_right = F.mvuRight[i] > 0;
            FrameRef df(ctx, 0, F, F.mnMinX, F.mnMinY, F.mnMaxX, F.mnMaxY, any_right);
            const int M = (int)vpMapPoints.size();
            std::vector<uint8_t> desc((size_t)M * 32), in_view(M), bad(M);
            std::vector<float> proj((size_t)M * 2), xr(M), cosv(M), depth(M);
            std::vector<int32_t> level(M), nobs(M);
            for (int i = 0; i < M; i++)
            {
                MapPointT *p = vpMapPoints[i];
                in_view[i] = p->mbTrackInView ? 1 : 0;
                bad[i] = p->isBad() ? 1 : 0;
                proj[2 * i] = p->mTrackProjX; proj[2 * i + 1] = p->mTrackProjY;
                xr[i] = p->mTrackProjXR; cosv[i] = p->mTrackViewCos; depth[i] = p->mTrackDepth;
                level[i] = p->mnTrackScaleLevel; nobs[i] = p->Observations();
                std::memcpy(&desc[(size_t)i * 32], p->GetDescriptor().template ptr<uint8_t>(), 32);
            }
            orbgpu_mappoints_host mh;
            std::memset(&mh, 0, sizeof(mh));
            mh.n = M; mh.desc = desc.data(); mh.proj_xy = proj.data(); mh.proj_xr = xr.data(); mh.scale_level = level.data();
            mh.view_cos = cosv.data(); mh.depth = depth.data(); mh.in_view = in_view.data(); mh.bad = bad.data(); mh.n_obs = nobs.data();
            std::vector<int32_t> prior(F.N > 0 ? F.N : 1, 0), kp_mp(F.N > 0 ? F.N : 1, -1);
            for (int i = 0; i < F.N; i++)
                if (F.mvpMapPoints[i]) prior[i] = F.mvpMapPoints[i]->Observations();
            int32_t nmatches = 0;
            check(orbgpu_search_by_projection_local(ctx, df.f, &mh, th, bFarPoints ? 1 : 0, thFarPoints, mfNNratio, prior.data(),
                                                    kp_mp.data(), &nmatches));
            for (int i = 0; i < F.N; i++)
                if (kp_mp[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[kp_mp[i]]; // :156
            return nmatches;
        }

        // Tracking::SearchLocalPoints (Tracking.cc:4125-4182) in one device call: Frame::isInFrustum(pMP, viewingCosLimit) of every local
        // map point the loop tests -- not already seen by this frame (mnLastFrameSeen == F.mnId, :4133), not bad (:4136) -- then
        // SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) on the projections, which never leave HBM.  Per point the
        // adapter writes mbTrackInView (what the caller's IncreaseVisible() loop and the matcher's :55 test read); *nToMatch = number
        // of points in view (:4140-4144).  The pose members are read exactly as Frame::UpdatePoseMatrices fills them (Frame.cc:
        // mRcw = mTcw.rotationMatrix(), mtcw = mTcw.translation(), mOw = GetCameraCenter()).  MapPointT needs GetMinDistanceRaw() /
        // GetMaxDistanceRaw(): PredictScale reads the raw mfMaxDistance.  With frustum_report().enabled the
        // reference's own host isInFrustum is run as the checker and every disagreement is counted.
        int SearchLocalPoints(FrameT &F, const std::vector<MapPointT *> &vpMapPoints, const float th = 1, const bool bFarPoints = false,
                              const float thFarPoints = 50.0f, const float viewingCosLimit = 0.5f, int *nToMatch = nullptr)
        {
            if (F.Nleft != -1) throw std::runtime_error("orbmatch_b200: stereo-fisheye (Nleft != -1) path is not on the GPU hot path");
            orbgpu_ctx *ctx = thread_context();
            bool any_right = false;
            for (int i = 0; i < F.N && !any_right; i++) any_right = F.mvuRight[i] > 0;
            FrameRef df(ctx, 0, F, F.mnMinX, F.mnMinY, F.mnMaxX, F.mnMaxY, any_right);
            orbgpu_frustum_host fr;
            std::memset(&fr, 0, sizeof(fr));
            const auto Tcw = F.GetPose();
            const auto R = Tcw.rotationMatrix();
            const auto t = Tcw.translation();
            const auto Ow = F.GetCameraCenter();
            for (int r = 0; r < 3; r++)
            {
                for (int c = 0; c < 3; c++) fr.Rcw[3 * r + c] = R(r, c);
                fr.tcw[r] = t(r);
                fr.Ow[r] = Ow(r);
            }
            fr.K[0] = F.fx; fr.K[1] = F.fy; fr.K[2] = F.cx; fr.K[3] = F.cy;
            fr.mbf = F.mbf;
            fr.min_x = F.mnMinX; fr.min_y = F.mnMinY; fr.max_x = F.mnMaxX; fr.max_y = F.mnMaxY;
            fr.viewing_cos_limit = viewingCosLimit;
            fr.log_scale_factor = F.mfLogScaleFactor;
            fr.n_levels = F.mnScaleLevels;
            const int M = (int)vpMapPoints.size();
            std::vector<uint8_t> desc((size_t)M * 32, 0), skip(M, 1), bad(M, 0), in_view(M > 0 ? M : 1, 0);
            std::vector<float> wp((size_t)M * 3, 0.f), nm((size_t)M * 3, 0.f), mind(M, 0.f), maxd(M, 0.f);
            std::vector<int32_t> nobs(M, 0);
            for (int i = 0; i < M; i++)
            {
                MapPointT *p = vpMapPoints[i];
                if (!p) continue;
                bad[i] = p->isBad() ? 1 : 0;
                if (p->mnLastFrameSeen == F.mnId || bad[i]) { p->mbTrackInView = p->mnLastFrameSeen == F.mnId ? false : p->mbTrackInView; continue; }
                skip[i] = 0;
                const auto P = p->GetWorldPos();
                const auto Pn = p->GetNormal();
                for (int c = 0; c < 3; c++) { wp[3 * i + c] = P(c); nm[3 * i + c] = Pn(c); }
                mind[i] = p->GetMinDistanceRaw();
                maxd[i] = p->GetMaxDistanceRaw();
                nobs[i] = p->Observations();
                std::memcpy(&desc[(size_t)i * 32], p->GetDescriptor().template ptr<uint8_t>(), 32);
            }
            orbgpu_localpoints_host lp;
            std::memset(&lp, 0, sizeof(lp));
            lp.n = M; lp.desc = desc.data(); lp.world_pos = wp.data(); lp.normal = nm.data(); lp.min_distance = mind.data();
            lp.max_distance = maxd.data(); lp.skip = skip.data(); lp.bad = bad.data(); lp.n_obs = nobs.data();
            std::vector<int32_t> prior(F.N > 0 ? F.N : 1, 0), kp_mp(F.N > 0 ? F.N : 1, -1);
            for (int i = 0; i < F.N; i++)
                if (F.mvpMapPoints[i]) prior[i] = F.mvpMapPoints[i]->Observations();
            int32_t nmatches = 0;
            check(orbgpu_search_local_points(ctx, df.f, &fr, &lp, th, bFarPoints ? 1 : 0, thFarPoints, mfNNratio, prior.data(), kp_mp.data(),
                                             in_view.data(), &nmatches));
            int n_in_view = 0;
            FrustumReport &rep = frustum_report();
            std::vector<uint8_t> g_iv;
            std::vector<float> g_xy, g_xr, g_dp, g_vc;
            std::vector<int32_t> g_lv;
            if (rep.enabled && M > 0)
            {   // the device's full isInFrustum record, to compare member by member with the host's
                g_iv.resize(M); g_xy.resize((size_t)M * 2); g_xr.resize(M); g_dp.resize(M); g_vc.resize(M); g_lv.resize(M);
                check(orbgpu_is_in_frustum(ctx, &fr, M, wp.data(), nm.data(), mind.data(), maxd.data(), g_iv.data(), g_xy.data(), g_xr.data(),
                                           g_dp.data(), g_lv.data(), g_vc.data()));
            }
            for (int i = 0; i < M; i++)
            {
                MapPointT *p = vpMapPoints[i];
                if (!p || skip[i]) continue;
                if (rep.enabled)
                {
                    const bool host = F.isInFrustum(p, viewingCosLimit); // the reference's own code (writes the members like it always did)
                    rep.points++;
                    if (host != (in_view[i] != 0)) rep.in_view_mismatch++;
                    else if (host)
                    {
                        if (p->mnTrackScaleLevel != g_lv[i]) rep.level_mismatch++;
                        if (p->mTrackProjX != g_xy[2 * i] || p->mTrackProjY != g_xy[2 * i + 1] || p->mTrackProjXR != g_xr[i]) rep.projection_mismatch++;
                    }
                }
                p->mbTrackInView = in_view[i] != 0;
                n_in_view += in_view[i] != 0;
            }
            if (nToMatch) *nToMatch = n_in_view;
            for (int i = 0; i < F.N; i++)
                if (kp_mp[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[kp_mp[i]]; // ORBmatcher.cc:156
            return nmatches;
        }

        // ORBmatcher.h:65 (ORBmatcher.cc:262-496)
        int SearchByBoW(KeyFrameT *pKF, FrameT &F, std::vector<MapPointT *> &vpMapPointMatches)
        {
            orbgpu_ctx *ctx = thread_context();
            const std::vector<MapPointT *> vpMapPointsKF = pKF->GetMapPointMatches();
            FrameRef dk(ctx, 1, *pKF, (float)pKF->mnMinX, (float)pKF->mnMinY, (float)pKF->mnMaxX, (float)pKF->mnMaxY, false);
            FrameRef df(ctx, 0, F, F.mnMinX, F.mnMinY, F.mnMaxX, F.mnMaxY, false);
            std::vector<uint8_t> valid(pKF->N > 0 ? pKF->N : 1, 0);
            for (int i = 0; i < pKF->N; i++) valid[i] = (vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad()) ? 1 : 0; // :311-315
            std::vector<int32_t> m(F.N > 0 ? F.N : 1, -1);
            int32_t nmatches = 0;
            check(orbgpu_search_by_bow_kf_f(ctx, dk.f, df.f, valid.data(), mfNNratio, mbCheckOrientation ? 1 : 0, m.data(), &nmatches));
            vpMapPointMatches = std::vector<MapPointT *>(F.N, static_cast<MapPointT *>(NULL)); // :268
            for (int i = 0; i < F.N; i++)
                if (m[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[m[i]]; // :398
            return nmatches;
        }

        // ORBmatcher.h:66 (ORBmatcher.cc:890-1043)
        int SearchByBoW(KeyFrameT *pKF1, KeyFrameT *pKF2, std::vector<MapPointT *> &vpMatches12)
        {
            orbgpu_ctx *ctx = thread_context();
            const std::vector<MapPointT *> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
            FrameRef d1(ctx, 1, *pKF1, (float)pKF1->mnMinX, (float)pKF1->mnMinY, (float)pKF1->mnMaxX, (float)pKF1->mnMaxY, false);
            FrameRef d2(ctx, 1, *pKF2, (float)pKF2->mnMinX, (float)pKF2->mnMinY, (float)pKF2->mnMaxX, (float)pKF2->mnMaxY, false);
            std::vector<uint8_t> v1(pKF1->N > 0 ? pKF1->N : 1, 0), v2(pKF2->N > 0 ? pKF2->N : 1, 0);
            for (int i = 0; i < pKF1->N; i++) v1[i] = (vp1[i] && !vp1[i]->isBad()) ? 1 : 0;
            for (int i = 0; i < pKF2->N; i++) v2[i] = (vp2[i] && !vp2[i]->isBad()) ? 1 : 0;
            std::vector<int32_t> m(pKF1->N > 0 ? pKF1->N : 1, -1);
            int32_t nmatches = 0;
            check(orbgpu_search_by_bow_kf_kf(ctx, d1.f, d2.f, v1.data(), v2.data(), mfNNratio, mbCheckOrientation ? 1 : 0, m.data(), &nmatches));
            vpMatches12 = std::vector<MapPointT *>(vp1.size(), static_cast<MapPointT *>(NULL)); // :904
            for (int i = 0; i < pKF1->N; i++)
                if (m[i] >= 0) vpMatches12[i] = vp2[m[i]]; // :989
            return nmatches;
        }

        // Batched forms of the two SearchByBoW overloads (not in ORBmatcher.h: the reference calls them in a loop).  One device call
        // for the whole loop; entry i equals what the single call on vpKFs[i] returns.  NULL entries are skipped (0 matches).
        //   Tracking::Relocalization (Tracking.cc:4469-4495):  vnMatches[i] = SearchByBoW(vpKFs[i], F, vvpMapPointMatches[i])
        void SearchByBoW(const std::vector<KeyFrameT *> &vpKFs, FrameT &F, std::vector<std::vector<MapPointT *>> &vvpMapPointMatches,
                         std::vector<int> &vnMatches)
        {
            bow_batch(0, vpKFs, F, F.mnMinX, F.mnMinY, F.mnMaxX, F.mnMaxY, nullptr, vvpMapPointMatches, vnMatches);
        }
        //   LoopClosing's candidate window (LoopClosing.cc:909-925):  vnMatches[j] = SearchByBoW(pKF1, vpKF2s[j], vvpMatches12[j])
        void SearchByBoW(KeyFrameT *pKF1, const std::vector<KeyFrameT *> &vpKF2s, std::vector<std::vector<MapPointT *>> &vvpMatches12,
                         std::vector<int> &vnMatches)
        {
            const std::vector<MapPointT *> vp1 = pKF1->GetMapPointMatches();
            bow_batch(1, vpKF2s, *pKF1, (float)pKF1->mnMinX, (float)pKF1->mnMinY, (float)pKF1->mnMaxX, (float)pKF1->mnMaxY, &vp1, vvpMatches12,
                      vnMatches);
        }

        // ORBmatcher.h:72 (ORBmatcher.cc:1045-1328), monocular pinhole path (mpCamera2 == NULL).
        // The pose algebra (epipole, R12, t12, F12) is evaluated HERE with the host's own Sophus / Eigen, exactly as the
        // reference does at :1053-1071 and Pinhole.cpp:194-197, and handed to the GPU as inputs.
        template <class Vec3 = void>
        int SearchForTriangulation(KeyFrameT *pKF1, KeyFrameT *pKF2, std::vector<std::pair<size_t, size_t>> &vMatchedPairs,
                                   const bool bOnlyStereo, const bool bCoarse = false)
        {
            if (pKF1->mpCamera2 || pKF2->mpCamera2) throw std::runtime_error("orbmatch_b200: two-camera rigs are not on the GPU hot path");
            orbgpu_ctx *ctx = thread_context();
            auto T1w = pKF1->GetPose();
            auto T2w = pKF2->GetPose();
            auto Tw2 = pKF2->GetPoseInverse();
            auto Cw = pKF1->GetCameraCenter();
            auto C2 = T2w * Cw;
            auto epv = pKF2->mpCamera->project(C2);
            auto T12 = T1w * Tw2;
            auto R12 = T12.rotationMatrix();
            auto t12 = T12.translation();
            // Pinhole.cpp:194-197
            auto t12x = hat(t12, R12);
            auto K1 = pKF1->mpCamera->toK_();
            auto K2 = pKF2->mpCamera->toK_();
            auto F12 = K1.transpose().inverse() * t12x * R12 * K2.inverse();
            float ep[2] = {epv(0), epv(1)}, f12[9];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) f12[3 * r + c] = F12(r, c);

            // a kfset has a uniform feature count: the smaller keyframe is padded with features that carry a map
            // point (has_mp = 1) and no vocabulary node, which the search skips (:1129, :1165)
            const int n = pKF1->N > pKF2->N ? pKF1->N : pKF2->N;
            std::vector<uint8_t> desc((size_t)2 * n * 32, 0), has_mp((size_t)2 * n, 1);
            std::vector<float> xy((size_t)4 * n), angle((size_t)2 * n), ur((size_t)2 * n);
            std::vector<int32_t> octave((size_t)2 * n);
            std::vector<uint32_t> node((size_t)2 * n, 0xFFFFFFFFu);
            KeyFrameT *kfs[2] = {pKF1, pKF2};
            bool any_right = false;
            for (int k = 0; k < 2; k++)
            {
                KeyFrameT *K = kfs[k];
                for (int i = 0; i < K->N; i++)
                {
                    const size_t o = (size_t)k * n + i;
                    std::memcpy(&desc[o * 32], K->mDescriptors.row(i).template ptr<uint8_t>(), 32);
                    xy[2 * o] = K->mvKeysUn[i].pt.x; xy[2 * o + 1] = K->mvKeysUn[i].pt.y;
                    octave[o] = K->mvKeysUn[i].octave; angle[o] = K->mvKeysUn[i].angle;
                    has_mp[o] = K->GetMapPoint(i) ? 1 : 0;
                    ur[o] = K->mvuRight[i];
                    any_right = any_right || K->mvuRight[i] >= 0;
                }
                for (const auto &kv : K->mFeatVec)
                    for (unsigned f : kv.second) node[(size_t)k * n + f] = kv.first;
            }
            std::vector<float> sf(pKF2->mvScaleFactors.begin(), pKF2->mvScaleFactors.end()), s2(pKF2->mvLevelSigma2.begin(), pKF2->mvLevelSigma2.end());
            orbgpu_kfset_host sh;
            std::memset(&sh, 0, sizeof(sh));
            sh.n_kf = 2; sh.n_feat = n; sh.desc = desc.data(); sh.kp_xy = xy.data(); sh.octave = octave.data(); sh.angle = angle.data();
            sh.has_mp = has_mp.data(); sh.u_right = any_right ? ur.data() : nullptr; sh.node_id = node.data();
            sh.n_levels = (int32_t)sf.size(); sh.scale_factors = sf.data(); sh.level_sigma2 = s2.data();
            orbgpu_kfset *set = nullptr;
            check(orbgpu_kfset_upload(ctx, &sh, &set));
            const int32_t k1 = 0, k2 = 1;
            std::vector<int32_t> m(n > 0 ? n : 1, -1);
            int32_t nmatches = 0;
            const int rc = orbgpu_search_for_triangulation_batch(ctx, set, 1, &k1, &k2, ep, f12, bOnlyStereo ? 1 : 0, bCoarse ? 1 : 0,
                                                                 mbCheckOrientation ? 1 : 0, m.data(), &nmatches);
            orbgpu_kfset_destroy(set);
            check(rc);
            vMatchedPairs.clear(); // :1317-1325
            vMatchedPairs.reserve(nmatches);
            for (int i = 0; i < pKF1->N; i++)
                if (m[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m[i]));
            return nmatches;
        }

        // ORBmatcher.h:48 (ORBmatcher.cc:1957-2191), monocular / RGB-D path (Nleft == -1)
        int SearchByProjection(FrameT &CurrentFrame, const FrameT &LastFrame, const float th, const bool bMono)
        {
            if (CurrentFrame.Nleft != -1 || LastFrame.Nleft != -1)
                throw std::runtime_error("orbmatch_b200: stereo-fisheye (Nleft != -1) path is not on the GPU hot path");
            orbgpu_ctx *ctx = thread_context();
            const auto Tcw = CurrentFrame.GetPose();
            const auto twc = Tcw.inverse().translation();
            const auto Tlw = LastFrame.GetPose();
            const auto tlc = Tlw * twc;
            const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;   // :1971
            const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono; // :1972
            ProjBatch b(LastFrame.N);
            for (int i = 0; i < LastFrame.N; i++)
            {
                MapPointT *pMP = LastFrame.mvpMapPoints[i];
                if (!pMP || LastFrame.mvbOutlier[i]) continue;
                const auto x3Dw = pMP->GetWorldPos();
                const auto x3Dc = Tcw * x3Dw;
                const float invzc = 1.0 / x3Dc(2); // :2005
                if (invzc < 0) continue;
                const auto uv = CurrentFrame.mpCamera->project(x3Dc);
                if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
                if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
                const int nLastOctave = LastFrame.mvKeys[i].octave;
                const float radius = th * CurrentFrame.mvScaleFactors[nLastOctave];
                int lo, hi; // :2031-2041
                if (bForward) { lo = nLastOctave; hi = -1; }
                else if (bBackward) { lo = 0; hi = nLastOctave; }
                else { lo = nLastOctave - 1; hi = nLastOctave + 1; }
                b.set(i, pMP->GetDescriptor(), uv(0), uv(1), radius, lo, hi);
                b.ur[i] = uv(0) - CurrentFrame.mbf * invzc; // :2056
                b.locks[i] = pMP->Observations() > 0 ? 1 : 0; // what a later point's :2046-2049 test will see
                b.angle[i] = LastFrame.mvKeysUn[i].angle;
            }
            bool any_right = false;
            for (int i = 0; i < CurrentFrame.N && !any_right; i++) any_right = CurrentFrame.mvuRight[i] > 0;
            FrameRef df(ctx, 0, CurrentFrame, CurrentFrame.mnMinX, CurrentFrame.mnMinY, CurrentFrame.mnMaxX, CurrentFrame.mnMaxY, any_right);
            std::vector<uint8_t> locked(CurrentFrame.N > 0 ? CurrentFrame.N : 1, 0);
            for (int i = 0; i < CurrentFrame.N; i++)
                locked[i] = (CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0) ? 1 : 0;
            const ProjResult r = run_projected(ctx, df.f, CurrentFrame.N, b, (float)ORBGPU_TH_HIGH, true, any_right, false,
                                               mbCheckOrientation, nullptr, &locked);
            scatter_owners(r, CurrentFrame.mvpMapPoints, [&](int m) { return LastFrame.mvpMapPoints[m]; });
            return r.nmatches;
        }

        // ORBmatcher.h:52 (ORBmatcher.cc:2203-2330)
        int SearchByProjection(FrameT &CurrentFrame, KeyFrameT *pKF, const std::set<MapPointT *> &sAlreadyFound, const float th,
                               const int ORBdist)
        {
            orbgpu_ctx *ctx = thread_context();
            const auto Tcw = CurrentFrame.GetPose();
            const auto Ow = Tcw.inverse().translation();
            const std::vector<MapPointT *> vpMPs = pKF->GetMapPointMatches();
            ProjBatch b((int)vpMPs.size());
            for (size_t i = 0; i < vpMPs.size(); i++)
            {
                MapPointT *pMP = vpMPs[i];
                if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
                const auto x3Dw = pMP->GetWorldPos();
                const auto x3Dc = Tcw * x3Dw;
                const auto uv = CurrentFrame.mpCamera->project(x3Dc);
                if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
                if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
                const auto PO = x3Dw - Ow;
                const float dist3D = PO.norm();
                const float maxDistance = pMP->GetMaxDistanceInvariance();
                const float minDistance = pMP->GetMinDistanceInvariance();
                if (dist3D < minDistance || dist3D > maxDistance) continue;
                const int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
                const float radius = th * CurrentFrame.mvScaleFactors[nPredictedLevel];
                b.set((int)i, pMP->GetDescriptor(), uv(0), uv(1), radius, nPredictedLevel - 1, nPredictedLevel + 1);
                b.angle[i] = pKF->mvKeysUn[i].angle;
            }
            FrameRef df(ctx, 0, CurrentFrame, CurrentFrame.mnMinX, CurrentFrame.mnMinY, CurrentFrame.mnMaxX, CurrentFrame.mnMaxY, false);
            std::vector<uint8_t> locked(CurrentFrame.N > 0 ? CurrentFrame.N : 1, 0);
            for (int i = 0; i < CurrentFrame.N; i++) locked[i] = CurrentFrame.mvpMapPoints[i] ? 1 : 0; // :2262-2263
            const ProjResult r = run_projected(ctx, df.f, CurrentFrame.N, b, (float)ORBdist, true, false, false, mbCheckOrientation, nullptr,
                                               &locked);
            scatter_owners(r, CurrentFrame.mvpMapPoints, [&](int m) { return vpMPs[m]; });
            return r.nmatches;
        }

        // ORBmatcher.h:56 (ORBmatcher.cc:498-620)
        template <class Sim3T>
        int SearchByProjection(KeyFrameT *pKF, Sim3T &Scw, const std::vector<MapPointT *> &vpPoints, std::vector<MapPointT *> &vpMatched,
                               int th, float ratioHamming = 1.0)
        {
            const ProjResult r = sim3_projection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming);
            for (int k = 0; k < pKF->N; k++)
                if (r.kp_owner[k] >= 0) vpMatched[k] = vpPoints[r.kp_owner[k]]; // :614
            return r.nmatches;
        }

        // ORBmatcher.h:60 (ORBmatcher.cc:622-733): same search, also records the keyframe each point came from
        template <class Sim3T>
        int SearchByProjection(KeyFrameT *pKF, Sim3T &Scw, const std::vector<MapPointT *> &vpPoints,
                               const std::vector<KeyFrameT *> &vpPointsKFs, std::vector<MapPointT *> &vpMatched,
                               std::vector<KeyFrameT *> &vpMatchedKF, int th, float ratioHamming = 1.0)
        {
            const ProjResult r = sim3_projection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming, true);
            for (int k = 0; k < pKF->N; k++)
                if (r.kp_owner[k] >= 0)
                {
                    vpMatched[k] = vpPoints[r.kp_owner[k]];       // :726
                    vpMatchedKF[k] = vpPointsKFs[r.kp_owner[k]];  // :727
                }
            return r.nmatches;
        }

        // ORBmatcher.h:81 (ORBmatcher.cc:1330-1545), left camera (bRight == false)
        int Fuse(KeyFrameT *pKF, const std::vector<MapPointT *> &vpMapPoints, const float th = 3.0, const bool bRight = false)
        {
            if (bRight) throw std::runtime_error("orbmatch_b200: two-camera rigs (bRight) are not on the GPU hot path");
            orbgpu_ctx *ctx = thread_context();
            const auto Tcw = pKF->GetPose();
            const auto Ow = pKF->GetCameraCenter();
            const float bf = pKF->mbf;
            const int nMPs = (int)vpMapPoints.size();
            ProjBatch b(nMPs);
            for (int i = 0; i < nMPs; i++)
            {
                MapPointT *pMP = vpMapPoints[i];
                if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
                const auto p3Dw = pMP->GetWorldPos();
                const auto p3Dc = Tcw * p3Dw;
                if (p3Dc(2) < 0.0f) continue;
                const float invz = 1 / p3Dc(2);
                const auto uv = pKF->mpCamera->project(p3Dc);
                if (!pKF->IsInImage(uv(0), uv(1))) continue;
                const float ur = uv(0) - bf * invz;
                const float maxDistance = pMP->GetMaxDistanceInvariance();
                const float minDistance = pMP->GetMinDistanceInvariance();
                const auto PO = p3Dw - Ow;
                const float dist3D = PO.norm();
                if (dist3D < minDistance || dist3D > maxDistance) continue;
                const auto Pn = pMP->GetNormal();
                if (PO.dot(Pn) < 0.5 * dist3D) continue;
                const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
                const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
                b.set(i, pMP->GetDescriptor(), uv(0), uv(1), radius, nPredictedLevel - 1, nPredictedLevel);
                b.ur[i] = ur;
            }
            bool any_right = false;
            for (int i = 0; i < pKF->N && !any_right; i++) any_right = pKF->mvuRight[i] >= 0;
            FrameRef dk(ctx, 1, *pKF, (float)pKF->mnMinX, (float)pKF->mnMinY, (float)pKF->mnMaxX, (float)pKF->mnMaxY, any_right);
            const std::vector<float> inv(pKF->mvInvLevelSigma2.begin(), pKF->mvInvLevelSigma2.end());
            const ProjResult r = run_projected(ctx, dk.f, pKF->N, b, (float)ORBGPU_TH_LOW, false, false, true, false, &inv, nullptr);
            // the map mutations stay on the host, applied in the reference's order (:1502-1523)
            int nFused = 0;
            for (int i = 0; i < nMPs; i++)
            {
                const int bestIdx = r.best_idx[i];
                if (bestIdx < 0) continue;
                MapPointT *pMP = vpMapPoints[i];
                if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue; // an earlier entry of the list may have changed it
                MapPointT *pMPinKF = pKF->GetMapPoint(bestIdx);
                if (pMPinKF)
                {
                    if (!pMPinKF->isBad())
                    {
                        if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                        else pMPinKF->Replace(pMP);
                    }
                }
                else
                {
                    pMP->AddObservation(pKF, bestIdx);
                    pKF->AddMapPoint(pMP, bestIdx);
                }
                nFused++;
            }
            return nFused;
        }

        // ORBmatcher.h:84 (ORBmatcher.cc:1547-1682)
        template <class Sim3T>
        int Fuse(KeyFrameT *pKF, Sim3T &Scw, const std::vector<MapPointT *> &vpPoints, float th, std::vector<MapPointT *> &vpReplacePoint)
        {
            orbgpu_ctx *ctx = thread_context();
            typedef decltype(pKF->GetPose()) SE3T;
            const SE3T Tcw = SE3T(Scw.rotationMatrix(), Scw.translation() / Scw.scale());
            const auto Ow = Tcw.inverse().translation();
            const std::set<MapPointT *> spAlreadyFound = pKF->GetMapPoints();
            const int nPoints = (int)vpPoints.size();
            ProjBatch b(nPoints);
            for (int iMP = 0; iMP < nPoints; iMP++)
            {
                MapPointT *pMP = vpPoints[iMP];
                if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
                float u, v, dist3D;
                if (!project_checked(pKF, Tcw, Ow, pMP, u, v, dist3D)) continue;
                const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
                const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
                b.set(iMP, pMP->GetDescriptor(), u, v, radius, nPredictedLevel - 1, nPredictedLevel);
            }
            FrameRef dk(ctx, 1, *pKF, (float)pKF->mnMinX, (float)pKF->mnMinY, (float)pKF->mnMaxX, (float)pKF->mnMaxY, false);
            const ProjResult r = run_projected(ctx, dk.f, pKF->N, b, (float)ORBGPU_TH_LOW, false, false, false, false, nullptr, nullptr);
            int nFused = 0;
            for (int iMP = 0; iMP < nPoints; iMP++) // :1657-1672
            {
                const int bestIdx = r.best_idx[iMP];
                if (bestIdx < 0) continue;
                MapPointT *pMP = vpPoints[iMP];
                MapPointT *pMPinKF = pKF->GetMapPoint(bestIdx);
                if (pMPinKF)
                {
                    if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
                }
                else
                {
                    pMP->AddObservation(pKF, bestIdx);
                    pKF->AddMapPoint(pMP, bestIdx);
                }
                nFused++;
            }
            return nFused;
        }

        // ORBmatcher.h:78 (ORBmatcher.cc:1684-1955)
        template <class Sim3T>
        int SearchBySim3(KeyFrameT *pKF1, KeyFrameT *pKF2, std::vector<MapPointT *> &vpMatches12, const Sim3T &S12, const float th)
        {
            orbgpu_ctx *ctx = thread_context();
            const float fx = pKF1->fx, fy = pKF1->fy, cx = pKF1->cx, cy = pKF1->cy;
            const auto T1w = pKF1->GetPose();
            const auto T2w = pKF2->GetPose();
            const Sim3T S21 = S12.inverse();
            const std::vector<MapPointT *> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
            const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
            std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
            for (int i = 0; i < N1; i++) // :1711-1722
            {
                MapPointT *pMP = vpMatches12[i];
                if (!pMP) continue;
                vbAlreadyMatched1[i] = true;
                const int idx2 = std::get<0>(pMP->GetIndexInKeyFrame(pKF2));
                if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
            }
            // one direction: the map points of `from` (already transformed into its camera by Tfw) through S into `to`
            auto direction = [&](KeyFrameT *to, const std::vector<MapPointT *> &pts, const std::vector<bool> &done, const decltype(T1w) &Tfw,
                                 const Sim3T &S, int n_from) {
                ProjBatch b(n_from);
                for (int i = 0; i < n_from; i++)
                {
                    MapPointT *pMP = pts[i];
                    if (!pMP || done[i] || pMP->isBad()) continue;
                    const auto p3Dw = pMP->GetWorldPos();
                    const auto p3Dfrom = Tfw * p3Dw;
                    const auto p3Dto = S * p3Dfrom;
                    if (p3Dto(2) < 0.0) continue;
                    const float invz = 1.0 / p3Dto(2);
                    const float x = p3Dto(0) * invz, y = p3Dto(1) * invz;
                    const float u = fx * x + cx, v = fy * y + cy;
                    if (!to->IsInImage(u, v)) continue;
                    const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
                    const float dist3D = p3Dto.norm();
                    if (dist3D < minDistance || dist3D > maxDistance) continue;
                    const int nPredictedLevel = pMP->PredictScale(dist3D, to);
                    const float radius = th * to->mvScaleFactors[nPredictedLevel];
                    b.set(i, pMP->GetDescriptor(), u, v, radius, nPredictedLevel - 1, nPredictedLevel);
                }
                FrameRef dk(ctx, 1, *to, (float)to->mnMinX, (float)to->mnMinY, (float)to->mnMaxX, (float)to->mnMaxY, false);
                return run_projected(ctx, dk.f, to->N, b, (float)ORBGPU_TH_HIGH, false, false, false, false, nullptr, nullptr).best_idx;
            };
            const std::vector<int32_t> vnMatch1 = direction(pKF2, vpMapPoints1, vbAlreadyMatched1, T1w, S21, N1);
            const std::vector<int32_t> vnMatch2 = direction(pKF1, vpMapPoints2, vbAlreadyMatched2, T2w, S12, N2);
            int nFound = 0; // :1936-1950
            for (int i1 = 0; i1 < N1; i1++)
            {
                const int idx2 = vnMatch1[i1];
                if (idx2 >= 0 && vnMatch2[idx2] == i1)
                {
                    vpMatches12[i1] = vpMapPoints2[idx2];
                    nFound++;
                }
            }
            return nFound;
        }

    protected:
        // mode 0: K key frames against the frame `fixed` (results indexed by the frame's features, values = the key frames' map points);
        // mode 1: the key frame `fixed` (map points *vpFixed) against K key frames (indexed by its features, values = theirs)
        template <class FS>
        void bow_batch(int mode, const std::vector<KeyFrameT *> &vpKFs, FS &fixed, float minX, float minY, float maxX, float maxY,
                       const std::vector<MapPointT *> *vpFixed, std::vector<std::vector<MapPointT *>> &vvpMatches, std::vector<int> &vnMatches)
        {
            orbgpu_ctx *ctx = thread_context();
            const size_t K_all = vpKFs.size();
            vvpMatches.assign(K_all, std::vector<MapPointT *>());
            vnMatches.assign(K_all, 0);
            std::vector<size_t> live;
            for (size_t i = 0; i < K_all; i++)
                if (vpKFs[i]) live.push_back(i);
            const int K = (int)live.size(), n_out = fixed.N;
            if (K == 0) return;
            FrameRef dfix(ctx, mode == 0 ? 0 : 1, fixed, minX, minY, maxX, maxY, false);
            std::vector<std::unique_ptr<FrameRef>> refs;
            std::vector<const orbgpu_frame *> handles(K);
            std::vector<std::vector<MapPointT *>> vps(K);
            std::vector<std::vector<uint8_t>> valid(K);
            std::vector<const uint8_t *> valid_ptr(K);
            for (int k = 0; k < K; k++)
            {
                KeyFrameT *pKF = vpKFs[live[k]];
                vps[k] = pKF->GetMapPointMatches();
                refs.emplace_back(new FrameRef(ctx, 1, *pKF, (float)pKF->mnMinX, (float)pKF->mnMinY, (float)pKF->mnMaxX, (float)pKF->mnMaxY, false));
                handles[k] = refs.back()->f;
                valid[k].assign(pKF->N > 0 ? pKF->N : 1, 0);
                for (int i = 0; i < pKF->N; i++) valid[k][i] = (vps[k][i] && !vps[k][i]->isBad()) ? 1 : 0; // :311-315 / :941-957
                valid_ptr[k] = valid[k].data();
            }
            std::vector<int32_t> m((size_t)K * (n_out > 0 ? n_out : 1), -1), nm(K, 0);
            if (mode == 0)
                check(orbgpu_search_by_bow_kf_f_batch(ctx, K, handles.data(), dfix.f, valid_ptr.data(), mfNNratio, mbCheckOrientation ? 1 : 0,
                                                      m.data(), nm.data()));
            else
            {
                std::vector<uint8_t> v1(n_out > 0 ? n_out : 1, 0);
                for (int i = 0; i < n_out; i++) v1[i] = ((*vpFixed)[i] && !(*vpFixed)[i]->isBad()) ? 1 : 0;
                check(orbgpu_search_by_bow_kf_kf_batch(ctx, dfix.f, v1.data(), K, handles.data(), valid_ptr.data(), mfNNratio,
                                                       mbCheckOrientation ? 1 : 0, m.data(), nm.data()));
            }
            for (int k = 0; k < K; k++)
            {
                std::vector<MapPointT *> &out = vvpMatches[live[k]];
                out.assign(mode == 0 ? (size_t)n_out : vpFixed->size(), static_cast<MapPointT *>(NULL)); // :268 / :904
                for (int i = 0; i < n_out; i++)
                    if (m[(size_t)k * n_out + i] >= 0) out[i] = vps[k][m[(size_t)k * n_out + i]]; // :398 / :989
                vnMatches[live[k]] = nm[k];
            }
        }

        // writes the GPU's owner table into the caller's map-point slots: a keypoint some point took ends up holding the
        // last taker, or NULL when the rotation histogram culled it (ORBmatcher.cc:2074 + :2163-2186, :2290 + :2310-2325)
        template <class Get>
        static void scatter_owners(const ProjResult &r, std::vector<MapPointT *> &slots, Get &&point_of)
        {
            for (size_t m = 0; m < r.best_idx.size(); m++)
                if (r.best_idx[m] >= 0) slots[r.best_idx[m]] = static_cast<MapPointT *>(NULL);
            for (size_t k = 0; k < slots.size(); k++)
                if (r.kp_owner[k] >= 0) slots[k] = point_of(r.kp_owner[k]);
        }

        // prologue shared by the Sim3 searches (ORBmatcher.cc:524-557, :646-684, :1581-1611): false when a gate drops the point
        // inline_pinhole: the 8-argument overload projects with fx*x*invz + cx written out (:647-652) instead of mpCamera->project
        template <class SE3T, class Vec3>
        static bool project_checked(KeyFrameT *pKF, const SE3T &Tcw, const Vec3 &Ow, MapPointT *pMP, float &u, float &v, float &dist,
                                    bool inline_pinhole = false)
        {
            const auto p3Dw = pMP->GetWorldPos();
            const auto p3Dc = Tcw * p3Dw;
            if (p3Dc(2) < 0.0) return false;
            auto uv = pKF->mpCamera->project(p3Dc);
            if (inline_pinhole)
            {
                const float invz = 1 / p3Dc(2);
                const float x = p3Dc(0) * invz, y = p3Dc(1) * invz;
                uv(0) = pKF->fx * x + pKF->cx;
                uv(1) = pKF->fy * y + pKF->cy;
            }
            if (!pKF->IsInImage(uv(0), uv(1))) return false;
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
            const auto PO = p3Dw - Ow;
            dist = PO.norm();
            if (dist < minDistance || dist > maxDistance) return false;
            const auto Pn = pMP->GetNormal();
            if (PO.dot(Pn) < 0.5 * dist) return false;
            u = uv(0); v = uv(1);
            return true;
        }

        template <class Sim3T>
        ProjResult sim3_projection(KeyFrameT *pKF, Sim3T &Scw, const std::vector<MapPointT *> &vpPoints,
                                   const std::vector<MapPointT *> &vpMatched, int th, float ratioHamming, bool inline_pinhole = false)
        {
            orbgpu_ctx *ctx = thread_context();
            typedef decltype(pKF->GetPose()) SE3T;
            const SE3T Tcw = SE3T(Scw.rotationMatrix(), Scw.translation() / Scw.scale()); // :507
            const auto Ow = Tcw.inverse().translation();
            std::set<MapPointT *> spAlreadyFound(vpMatched.begin(), vpMatched.end());
            spAlreadyFound.erase(static_cast<MapPointT *>(NULL));
            const int n = (int)vpPoints.size();
            ProjBatch b(n);
            for (int iMP = 0; iMP < n; iMP++)
            {
                MapPointT *pMP = vpPoints[iMP];
                if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
                float u, v, dist;
                if (!project_checked(pKF, Tcw, Ow, pMP, u, v, dist, inline_pinhole)) continue;
                const int nPredictedLevel = pMP->PredictScale(dist, pKF);
                const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
                b.set(iMP, pMP->GetDescriptor(), u, v, radius, nPredictedLevel - 1, nPredictedLevel);
            }
            FrameRef dk(ctx, 1, *pKF, (float)pKF->mnMinX, (float)pKF->mnMinY, (float)pKF->mnMaxX, (float)pKF->mnMaxY, false);
            std::vector<uint8_t> locked(pKF->N > 0 ? pKF->N : 1, 0);
            for (int k = 0; k < pKF->N; k++) locked[k] = vpMatched[k] ? 1 : 0; // :580-581
            return run_projected(ctx, dk.f, pKF->N, b, ORBGPU_TH_LOW * ratioHamming, true, false, false, false, nullptr, &locked);
        }

        // Sophus::SO3f::hat(t12) written against the matrix type of R12 (so that no Sophus header is needed here)
        template <class V, class Mtx>
        static Mtx hat(const V &w, const Mtx &like)
        {
            Mtx O = like;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) O(r, c) = 0.f;
            O(0, 1) = -w(2); O(0, 2) = w(1);
            O(1, 0) = w(2);  O(1, 2) = -w(0);
            O(2, 0) = -w(1); O(2, 1) = w(0);
            return O;
        }

        float mfNNratio;
        bool mbCheckOrientation;
    };
} // namespace orbgpu
