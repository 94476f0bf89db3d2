// adapter_parity.cc -- C++ drop-in parity test (test infrastructure).
//
// Links the reference's OWN ORBmatcher.cc (compiled unmodified against the oracle's stub Frame /
// KeyFrame / MapPoint, oracle/shim) and runs it side by side with orbgpu::ORBmatcherT (the GPU
// adapter with the reference's call signatures) on the SAME C++ objects; every STL output must be
// identical.  Built by `make -C oracle adapter` where /root/reference is mounted; needs a B200 to run.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "ORBmatcher.h" // the reference's header (stubs resolve MapPoint.h / KeyFrame.h / Frame.h)
#include "orbmatch_b200/ORBmatcher.hpp"

// the reference's vocabulary (DBoW2, compiled unmodified) and the drop-in wrapper that overrides its transform
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
namespace std
{
    template <>
    struct pair<cv::Mat, DBoW2::FORB> // TemplatedVocabulary.h:435 declares a vector of this pair with abstract FORB (unused member)
    {
    };
}
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"
#include "Thirdparty/DBoW2/DUtils/Random.h"
#include "orbmatch_b200/ORBVocabulary.hpp"

using namespace ORB_SLAM3;
typedef orbgpu::ORBmatcherT<Frame, KeyFrame, MapPoint> GpuMatcher;

static std::mt19937_64 rng(20261018);
static float uni(float a, float b) { return a + (b - a) * (float)((rng() >> 11) * (1.0 / 9007199254740992.0)); }
static float quant(float v) { return std::round(v * 4.f) / 4.f; }

static void scale_tables(FeatureSet &s)
{
    s.mnScaleLevels = 8;
    s.mvScaleFactors.assign(8, 1.f);
    s.mvLevelSigma2.assign(8, 1.f);
    s.mvInvLevelSigma2.assign(8, 1.f);
    for (int i = 1; i < 8; i++)
    {
        s.mvScaleFactors[i] = s.mvScaleFactors[i - 1] * 1.2f;
        s.mvLevelSigma2[i] = s.mvScaleFactors[i] * s.mvScaleFactors[i];
        s.mvInvLevelSigma2[i] = 1.0f / s.mvLevelSigma2[i];
    }
    s.mfLogScaleFactor = std::log(1.2f);
    s.mfGridElementWidthInv = 64.f / 640.f;
    s.mfGridElementHeightInv = 48.f / 480.f;
}

static int octave()
{
    static const float share[8] = {0.217f, 0.181f, 0.151f, 0.126f, 0.105f, 0.087f, 0.073f, 0.060f};
    float u = uni(0, 1), acc = 0;
    for (int i = 0; i < 8; i++) { acc += share[i]; if (u < acc) return i; }
    return 7;
}

static void random_features(FeatureSet &s, int n)
{
    s.N = n;
    s.mvKeysUn.resize(n);
    s.mDescriptors.create(n, 32, CV_8U);
    for (int i = 0; i < n; i++)
    {
        cv::KeyPoint kp;
        kp.pt.x = quant(uni(0, 639.5f)); kp.pt.y = quant(uni(0, 479.5f));
        kp.octave = octave(); kp.angle = quant(uni(0, 359.5f));
        s.mvKeysUn[i] = kp;
        for (int b = 0; b < 32; b++) s.mDescriptors.ptr<uint8_t>(i)[b] = (uint8_t)(rng() & 0xFF);
    }
    s.mvKeys = s.mvKeysUn;
    s.mvuRight.assign(n, -1.f);
    s.mvpMapPoints.assign(n, nullptr);
    scale_tables(s);
}

// copy feature j of src into feature i of dst with ~flips random bit flips
static void plant(FeatureSet &dst, int i, const FeatureSet &src, int j, int flips, float dx, float dy, float drot)
{
    std::memcpy(dst.mDescriptors.ptr<uint8_t>(i), src.mDescriptors.ptr<uint8_t>(j), 32);
    for (int f = 0; f < flips; f++) { int bit = (int)(rng() % 256); dst.mDescriptors.ptr<uint8_t>(i)[bit >> 3] ^= (uint8_t)(1 << (bit & 7)); }
    cv::KeyPoint kp = src.mvKeysUn[j];
    kp.pt.x = quant(std::min(639.f, std::max(0.f, kp.pt.x + dx)));
    kp.pt.y = quant(std::min(479.f, std::max(0.f, kp.pt.y + dy)));
    kp.angle = quant(std::fmod(kp.angle + drot + 720.f, 360.f));
    dst.mvKeysUn[i] = kp;
    dst.mvKeys[i] = kp;
}

static void finish_frame(Frame &F) { F.mnMinX = 0; F.mnMinY = 0; F.mnMaxX = 640; F.mnMaxY = 480; F.mvbOutlier.assign(F.N, false); F.AssignFeaturesToGrid(); }
static void finish_kf(KeyFrame &K)
{
    Frame F; // a key frame takes the grid of the frame it is made from (KeyFrame.cc:66-82)
    F.N = K.N; F.mvKeysUn = K.mvKeysUn;
    F.mfGridElementWidthInv = K.mfGridElementWidthInv; F.mfGridElementHeightInv = K.mfGridElementHeightInv;
    F.mnMinX = 0; F.mnMinY = 0; F.mnMaxX = 640; F.mnMaxY = 480;
    F.AssignFeaturesToGrid();
    K.CopyGridFrom(F);
}


// ---- a scene with real geometry for the overloads that project on their own.  Built twice from the
// same seed: the reference runs on one copy, the GPU adapter on the other (Fuse mutates key frames and map points).
struct TagPoint : public MapPoint
{
    int tag = -1;
};
struct Scene
{
    Pinhole cam{368.05096f, 368.05399f, 317.11264f, 236.39537f};
    Frame Cur, Last;
    KeyFrame K, K2;
    std::vector<TagPoint> pts, held, held2, twins;
    std::vector<int> pair_m, pair_tgt, pair_j; // landmark m: keypoint tgt of Cur/K and keypoint j of K2
    std::vector<MapPoint *> vp;
    std::vector<KeyFrame *> vpKFs;
    Sophus::Sim3f Scw, S12;
};
static Eigen::Matrix3f rot_y(float a)
{
    Eigen::Matrix3f R = Eigen::Matrix3f::Identity();
    R(0, 0) = std::cos(a); R(0, 2) = std::sin(a); R(2, 0) = -std::sin(a); R(2, 2) = std::cos(a);
    return R;
}
static int tag_of(MapPoint *p) { return p ? static_cast<TagPoint *>(p)->tag : -1; }
static std::vector<int> tags(const std::vector<MapPoint *> &v)
{
    std::vector<int> t(v.size());
    for (size_t i = 0; i < v.size(); i++) t[i] = tag_of(v[i]);
    return t;
}
// world point that projects onto (u, v) at depth z in a camera with pose Tcw
static Eigen::Vector3f backproject(const Pinhole &cam, const Sophus::SE3f &Tcw, float u, float v, float z)
{
    Eigen::Vector3f Xc((u - cam.mvParameters[2]) / cam.mvParameters[0] * z, (v - cam.mvParameters[3]) / cam.mvParameters[1] * z, z);
    return Tcw.inverse() * Xc;
}
static void build_scene(Scene &S, uint64_t seed, bool stereo)
{
    rng.seed(seed);
    const int N = 2000, M = 3000;
    const Sophus::SE3f Tcw(rot_y(0.02f), Eigen::Vector3f(0.10f, -0.03f, 0.05f));
    S.Scw = Sophus::Sim3f(1.08f, rot_y(0.02f), Eigen::Vector3f(0.10f, -0.03f, 0.05f) * 1.08f); // same camera through a Sim3
    for (FeatureSet *f : {(FeatureSet *)&S.Cur, (FeatureSet *)&S.K, (FeatureSet *)&S.K2})
    {
        random_features(*f, N);
        f->mpCamera = &S.cam;
        f->fx = S.cam.mvParameters[0]; f->fy = S.cam.mvParameters[1]; f->cx = S.cam.mvParameters[2]; f->cy = S.cam.mvParameters[3];
        f->mbf = 40.f; f->mb = 0.11f;
        f->mTcw = Tcw;
    }
    S.K.mDescriptors = S.Cur.mDescriptors.clone(); S.K.mvKeysUn = S.Cur.mvKeysUn; S.K.mvKeys = S.Cur.mvKeys;
    if (stereo)
        for (int i = 0; i < N; i++)
            if (uni(0, 1) < 0.6f) { S.Cur.mvuRight[i] = quant(std::max(0.25f, S.Cur.mvKeysUn[i].pt.x - uni(2, 40))); S.K.mvuRight[i] = S.Cur.mvuRight[i]; }
    // second key frame for SearchBySim3: another pose, features planted from the same landmarks below
    S.K2.mTcw = Sophus::SE3f(rot_y(-0.03f), Eigen::Vector3f(-0.25f, 0.02f, 0.04f));
    S.pts.resize(M); S.vp.resize(M); S.vpKFs.resize(M);
    S.Last.N = M; S.Last.mvKeys.resize(M); S.Last.mvKeysUn.resize(M); S.Last.mvpMapPoints.assign(M, nullptr); S.Last.mvbOutlier.assign(M, false);
    S.Last.mTcw = Sophus::SE3f(rot_y(0.018f), Eigen::Vector3f(0.10f, -0.03f, stereo ? 0.45f : 0.04f)); // stereo: a clear forward motion
    scale_tables(S.Last);
    for (int i = 0; i < M; i++)
    {
        TagPoint &p = S.pts[i];
        p.tag = i;
        const bool planted = uni(0, 1) < 0.6f;
        const int tgt = (int)(rng() % N);
        const cv::KeyPoint &kp = S.Cur.mvKeysUn[tgt];
        const float u = planted ? kp.pt.x + uni(-2.5f, 2.5f) : uni(-20, 660), v = planted ? kp.pt.y + uni(-2.5f, 2.5f) : uni(-20, 500);
        const float z = uni(2.f, 12.f);
        p.worldPos_ = backproject(S.cam, Tcw, u, v, z);
        p.descriptor_.create(1, 32, CV_8U);
        for (int b = 0; b < 32; b++) p.descriptor_.ptr<uint8_t>()[b] = planted ? S.Cur.mDescriptors.ptr<uint8_t>(tgt)[b] : (uint8_t)(rng() & 0xFF);
        if (planted) for (int f = 0; f < (int)(rng() % 30); f++) { int bit = (int)(rng() % 256); p.descriptor_.ptr<uint8_t>()[bit >> 3] ^= (uint8_t)(1 << (bit & 7)); }
        const Eigen::Vector3f Ow = Tcw.inverse().translation(), PO = p.worldPos_ - Ow;
        const float dist = PO.norm();
        const int lvl = planted ? std::min(7, kp.octave + (int)(rng() % 2)) : octave();
        p.mfMaxDistance = dist * std::pow(1.2f, (float)lvl - uni(0.2f, 0.8f));
        p.mfMinDistance = p.mfMaxDistance / std::pow(1.2f, 7.f) * uni(0.5f, 1.0f);
        Eigen::Vector3f nrm = PO / dist;
        if (uni(0, 1) < 0.05f) nrm = Eigen::Vector3f(-nrm(0), nrm(1), -nrm(2)); // viewing-angle gate
        p.normal_ = nrm;
        p.nObs_ = uni(0, 1) < 0.12f ? 0 : 1 + (int)(rng() % 4);
        p.bad_ = uni(0, 1) < 0.03f;
        S.vp[i] = &p;
        S.vpKFs[i] = (i & 1) ? &S.K : &S.K2;
        cv::KeyPoint lk;
        lk.octave = lvl; lk.angle = quant(std::fmod((planted ? kp.angle : uni(0, 360)) + 25.f + uni(-6, 6) + 720.f, 360.f));
        if (planted && uni(0, 1) < 0.15f) lk.angle = quant(uni(0, 359.5f));
        S.Last.mvKeys[i] = lk; S.Last.mvKeysUn[i] = lk;
        if (uni(0, 1) < 0.9f) S.Last.mvpMapPoints[i] = &p;
        S.Last.mvbOutlier[i] = uni(0, 1) < 0.05f;
        // the same landmark seen by K2 (for SearchBySim3)
        if (planted && uni(0, 1) < 0.5f)
        {
            const Eigen::Vector2f u2 = S.cam.project(S.K2.mTcw * p.worldPos_);
            if (u2(0) > 1 && u2(0) < 638 && u2(1) > 1 && u2(1) < 478)
            {
                const int j = (int)(rng() % N);
                plant(S.K2, j, S.Cur, tgt, (int)(rng() % 30), 0, 0, 10.f);
                S.K2.mvKeysUn[j].pt.x = quant(u2(0) + uni(-1.5f, 1.5f)); S.K2.mvKeysUn[j].pt.y = quant(u2(1) + uni(-1.5f, 1.5f));
                S.K2.mvKeysUn[j].octave = kp.octave; S.K2.mvKeys[j] = S.K2.mvKeysUn[j];
                S.pair_m.push_back(i); S.pair_tgt.push_back(tgt); S.pair_j.push_back(j);
            }
        }
    }
    finish_frame(S.Cur); finish_kf(S.K); finish_kf(S.K2);
    // what the frames / key frames hold on entry
    S.held.resize(N); S.held2.resize(N);
    for (int i = 0; i < N; i++)
    {
        S.held[i].tag = 100000 + i; S.held2[i].tag = 200000 + i;
        S.held[i].nObs_ = (int)(rng() % 3); S.held2[i].nObs_ = (int)(rng() % 3);
        S.held[i].bad_ = uni(0, 1) < 0.05f;
        if (uni(0, 1) < 0.12f) { S.Cur.mvpMapPoints[i] = &S.held[i]; S.K.mvpMapPoints[i] = &S.held[i]; S.held[i].observations_[&S.K] = std::make_tuple(i, -1); }
    }
    // SearchBySim3 walks the map points of both key frames: give K the landmarks at their planted keypoints
    S.S12 = Sophus::Sim3f(1.0f, (S.K.mTcw * S.K2.mTcw.inverse()).rotationMatrix(), (S.K.mTcw * S.K2.mTcw.inverse()).translation());
}

static int fails = 0;
#define EXPECT(cond, what)                                            \
    do {                                                              \
        if (!(cond)) { std::printf("MISMATCH: %s\n", what); fails++; } \
        else std::printf("ok: %s\n", what);                           \
    } while (0)

int main()
{
    std::setvbuf(stdout, nullptr, _IONBF, 0); // a crash must not swallow the log
    // ---- SearchForInitialization
    {
        Frame F1, F2;
        random_features(F1, 1000);
        random_features(F2, 1000);
        for (int i = 0; i < 700; i++) plant(F2, (int)(rng() % 1000), F1, (int)(rng() % 1000), (int)(rng() % 40), uni(-60, 60), uni(-60, 60), -37.f + uni(-8, 8));
        finish_frame(F1); finish_frame(F2);
        std::vector<cv::Point2f> prevA(1000), prevB(1000);
        for (int i = 0; i < 1000; i++) prevA[i] = prevB[i] = F1.mvKeysUn[i].pt;
        std::vector<int> mA, mB;
        ORBmatcher ref(0.9f, true);
        GpuMatcher gpu(0.9f, true);
        const int nA = ref.SearchForInitialization(F1, F2, prevA, mA, 100);
        const int nB = gpu.SearchForInitialization(F1, F2, prevB, mB, 100);
        bool same = nA == nB && mA == mB;
        for (int i = 0; i < 1000; i++) same = same && prevA[i].x == prevB[i].x && prevA[i].y == prevB[i].y;
        std::printf("SearchForInitialization: ref %d gpu %d\n", nA, nB);
        EXPECT(same && nA > 20, "SearchForInitialization vnMatches12 / vbPrevMatched / return value");
    }
    // ---- SearchByProjection(Frame, MapPoints)
    {
        Frame FA;
        random_features(FA, 2000);
        finish_frame(FA);
        const int M = 5000;
        std::vector<MapPoint> pts(M);
        std::vector<MapPoint *> vp(M);
        for (int i = 0; i < M; i++)
        {
            MapPoint &p = pts[i];
            p.descriptor_.create(1, 32, CV_8U);
            const int tgt = (int)(rng() % 2000);
            const bool planted = uni(0, 1) < 0.35f;
            for (int b = 0; b < 32; b++) p.descriptor_.ptr<uint8_t>()[b] = planted ? FA.mDescriptors.ptr<uint8_t>(tgt)[b] : (uint8_t)(rng() & 0xFF);
            if (planted) for (int f = 0; f < (int)(rng() % 30); f++) { int bit = (int)(rng() % 256); p.descriptor_.ptr<uint8_t>()[bit >> 3] ^= (uint8_t)(1 << (bit & 7)); }
            p.mTrackProjX = planted ? quant(FA.mvKeysUn[tgt].pt.x + uni(-2, 2)) : quant(uni(0, 639));
            p.mTrackProjY = planted ? quant(FA.mvKeysUn[tgt].pt.y + uni(-2, 2)) : quant(uni(0, 479));
            p.mnTrackScaleLevel = planted ? std::min(7, FA.mvKeysUn[tgt].octave + (int)(rng() % 2)) : octave();
            p.mTrackViewCos = uni(0, 1) < 0.2f ? 0.9995f : uni(0.9f, 1.0f);
            p.mTrackDepth = uni(0.5f, 80.f);
            p.mbTrackInView = uni(0, 1) < 0.9f;
            p.bad_ = uni(0, 1) < 0.03f;
            p.nObs_ = uni(0, 1) < 0.1f ? 0 : 1 + (int)(rng() % 5);
            vp[i] = &p;
        }
        std::vector<MapPoint> prior(2000);
        Frame FB = FA;
        for (int i = 0; i < 2000; i++)
            if (uni(0, 1) < 0.1f) { prior[i].nObs_ = (int)(rng() % 4); FA.mvpMapPoints[i] = &prior[i]; FB.mvpMapPoints[i] = &prior[i]; }
        for (float th : {1.0f, 3.0f})
        {
            Frame A = FA, B = FB;
            ORBmatcher ref(0.8f, true);
            GpuMatcher gpu(0.8f, true);
            const int nA = ref.SearchByProjection(A, vp, th, true, 60.f);
            const int nB = gpu.SearchByProjection(B, vp, th, true, 60.f);
            std::printf("SearchByProjection th=%.0f: ref %d gpu %d\n", th, nA, nB);
            EXPECT(nA == nB && A.mvpMapPoints == B.mvpMapPoints && nA > 100, "SearchByProjection F.mvpMapPoints / return value");
        }
    }
    // ---- SearchByBoW (both) and SearchForTriangulation
    {
        KeyFrame K1, K2;
        Frame F;
        random_features(K1, 2000);
        random_features(K2, 2000);
        random_features(F, 2000);
        std::vector<int> node1(2000), node2(2000), nodeF(2000);
        for (int i = 0; i < 2000; i++) { node1[i] = 11 + (int)(rng() % 100); node2[i] = 11 + (int)(rng() % 100); nodeF[i] = 11 + (int)(rng() % 100); }
        // world landmarks seen by both keyframes (epipolar-consistent) -- camera 2 = small pure translation + tiny rotation
        Pinhole cam(368.05096f, 368.05399f, 317.11264f, 236.39537f);
        Eigen::Matrix3f R = Eigen::Matrix3f::Identity();
        const float a = 0.03f;
        R(0, 0) = std::cos(a); R(0, 2) = std::sin(a); R(2, 0) = -std::sin(a); R(2, 2) = std::cos(a);
        K1.mTcw = Sophus::SE3f(Eigen::Matrix3f::Identity(), Eigen::Vector3f(0.f, 0.f, 0.f));
        K2.mTcw = Sophus::SE3f(R, Eigen::Vector3f(0.3f, 0.05f, 0.02f));
        K1.mpCamera = &cam; K2.mpCamera = &cam;
        for (int i = 0; i < 1200; i++)
        {
            Eigen::Vector3f X(uni(-4, 4), uni(-3, 3), uni(2.5f, 12.f));
            Eigen::Vector2f u1 = cam.project(K1.mTcw * X), u2 = cam.project(K2.mTcw * X);
            if (u1(0) < 0 || u1(0) > 639 || u1(1) < 0 || u1(1) > 479 || u2(0) < 0 || u2(0) > 639 || u2(1) < 0 || u2(1) > 479) continue;
            const int i1 = (int)(rng() % 2000), i2 = (int)(rng() % 2000);
            K1.mvKeysUn[i1].pt.x = quant(u1(0) + uni(-0.7f, 0.7f)); K1.mvKeysUn[i1].pt.y = quant(u1(1) + uni(-0.7f, 0.7f));
            plant(K2, i2, K1, i1, (int)(rng() % 35), 0, 0, 15.f + uni(-5, 5));
            K2.mvKeysUn[i2].pt.x = quant(u2(0) + uni(-0.7f, 0.7f)); K2.mvKeysUn[i2].pt.y = quant(u2(1) + uni(-0.7f, 0.7f));
            node2[i2] = node1[i1];
            const int iF = (int)(rng() % 2000);
            plant(F, iF, K1, i1, (int)(rng() % 35), uni(-5, 5), uni(-5, 5), -21.f + uni(-6, 6));
            nodeF[iF] = node1[i1];
        }
        K1.mvKeys = K1.mvKeysUn; K2.mvKeys = K2.mvKeysUn;
        for (int i = 0; i < 2000; i++) { K1.mFeatVec.addFeature(node1[i], i); K2.mFeatVec.addFeature(node2[i], i); F.mFeatVec.addFeature(nodeF[i], i); }
        finish_kf(K1); finish_kf(K2); finish_frame(F);
        std::vector<MapPoint> p1(2000), p2(2000);
        for (int i = 0; i < 2000; i++)
        {
            if (uni(0, 1) < 0.55f) K1.mvpMapPoints[i] = &p1[i];
            if (uni(0, 1) < 0.55f) K2.mvpMapPoints[i] = &p2[i];
            p1[i].bad_ = uni(0, 1) < 0.03f; p2[i].bad_ = uni(0, 1) < 0.03f;
        }
        {
            std::vector<MapPoint *> mA, mB;
            ORBmatcher ref(0.7f, true);
            GpuMatcher gpu(0.7f, true);
            const int nA = ref.SearchByBoW(&K1, F, mA), nB = gpu.SearchByBoW(&K1, F, mB);
            std::printf("SearchByBoW KF-F: ref %d gpu %d\n", nA, nB);
            EXPECT(nA == nB && mA == mB && nA > 50, "SearchByBoW(KeyFrame*, Frame&) vpMapPointMatches / return value");
        }
        {
            std::vector<MapPoint *> mA, mB;
            ORBmatcher ref(0.9f, true);
            GpuMatcher gpu(0.9f, true);
            const int nA = ref.SearchByBoW(&K1, &K2, mA), nB = gpu.SearchByBoW(&K1, &K2, mB);
            std::printf("SearchByBoW KF-KF: ref %d gpu %d\n", nA, nB);
            EXPECT(nA == nB && mA == mB && nA > 20, "SearchByBoW(KeyFrame*, KeyFrame*) vpMatches12 / return value");
        }
        { // the batched overloads == the reference's own loops (Tracking.cc:4469-4495, LoopClosing.cc:909-925), NULL entries skipped
            std::vector<KeyFrame *> cands = {&K1, static_cast<KeyFrame *>(NULL), &K2, &K1};
            ORBmatcher ref(0.75f, true);
            GpuMatcher gpu(0.75f, true);
            std::vector<std::vector<MapPoint *>> vvB;
            std::vector<int> vnB;
            gpu.SearchByBoW(cands, F, vvB, vnB);
            bool ok = vvB.size() == cands.size() && vnB.size() == cands.size();
            int total = 0;
            for (size_t i = 0; ok && i < cands.size(); i++)
            {
                if (!cands[i]) { ok = vnB[i] == 0 && vvB[i].empty(); continue; }
                std::vector<MapPoint *> mA;
                const int nA = ref.SearchByBoW(cands[i], F, mA);
                ok = nA == vnB[i] && mA == vvB[i];
                total += nA;
            }
            std::printf("SearchByBoW batch KF-F: %d matches over %zu candidates\n", total, cands.size());
            EXPECT(ok && total > 100, "SearchByBoW(vector<KeyFrame*>, Frame&) == the loop over the reference's SearchByBoW");
            std::vector<KeyFrame *> window = {&K2, &K1, static_cast<KeyFrame *>(NULL), &K2};
            gpu.SearchByBoW(&K1, window, vvB, vnB);
            ok = vvB.size() == window.size();
            total = 0;
            for (size_t i = 0; ok && i < window.size(); i++)
            {
                if (!window[i]) { ok = vnB[i] == 0 && vvB[i].empty(); continue; }
                std::vector<MapPoint *> mA;
                const int nA = ref.SearchByBoW(&K1, window[i], mA);
                ok = nA == vnB[i] && mA == vvB[i];
                total += nA;
            }
            std::printf("SearchByBoW batch KF-KF: %d matches over %zu window key frames\n", total, window.size());
            EXPECT(ok && total > 50, "SearchByBoW(KeyFrame*, vector<KeyFrame*>) == the loop over the reference's SearchByBoW");
        }
        for (int ori = 0; ori < 2; ori++)
        {
            std::vector<std::pair<size_t, size_t>> vA, vB;
            ORBmatcher ref(0.6f, ori != 0);
            GpuMatcher gpu(0.6f, ori != 0);
            const int nA = ref.SearchForTriangulation(&K1, &K2, vA, false, false), nB = gpu.SearchForTriangulation(&K1, &K2, vB, false, false);
            std::printf("SearchForTriangulation checkOri=%d: ref %d gpu %d\n", ori, nA, nB);
            EXPECT(nA == nB && vA == vB && nA > 20, "SearchForTriangulation vMatchedPairs / return value");
        }
    }
    // ---- row a6: the overloads that project on their own, on real geometry
    for (int stereo = 0; stereo < 2; stereo++)
    {
        {
            Scene A, B;
            build_scene(A, 77 + stereo, stereo != 0); build_scene(B, 77 + stereo, stereo != 0);
            ORBmatcher ref(0.9f, true);
            GpuMatcher gpu(0.9f, true);
            const float th = stereo ? 15.f : 7.f;
            const int nA = ref.SearchByProjection(A.Cur, A.Last, th, !stereo), nB = gpu.SearchByProjection(B.Cur, B.Last, th, !stereo);
            std::printf("SearchByProjection(Cur, Last) stereo=%d: ref %d gpu %d\n", stereo, nA, nB);
            EXPECT(nA == nB && tags(A.Cur.mvpMapPoints) == tags(B.Cur.mvpMapPoints) && nA > 100, "SearchByProjection(Frame&, const Frame&) mvpMapPoints / return value");
        }
        {
            Scene A, B;
            build_scene(A, 79 + stereo, false); build_scene(B, 79 + stereo, false);
            for (Scene *S : {&A, &B})
                for (int i = 0; i < 2000; i++) S->K.mvpMapPoints[i] = i < (int)S->vp.size() && (i % 3) ? S->vp[i] : nullptr; // the KF's map points
            std::set<MapPoint *> fA, fB;
            for (int i = 0; i < 2000; i += 7) { fA.insert(A.vp[i]); fB.insert(B.vp[i]); }
            ORBmatcher ref(0.9f, stereo != 0);
            GpuMatcher gpu(0.9f, stereo != 0);
            const int nA = ref.SearchByProjection(A.Cur, &A.K, fA, 10.f, 64), nB = gpu.SearchByProjection(B.Cur, &B.K, fB, 10.f, 64);
            std::printf("SearchByProjection(Cur, KF, found) ori=%d: ref %d gpu %d\n", stereo, nA, nB);
            EXPECT(nA == nB && tags(A.Cur.mvpMapPoints) == tags(B.Cur.mvpMapPoints) && nA > 100, "SearchByProjection(Frame&, KeyFrame*, set) mvpMapPoints / return value");
        }
    }
    {
        Scene A, B;
        build_scene(A, 81, false); build_scene(B, 81, false);
        std::vector<MapPoint *> mA(2000, nullptr), mB(2000, nullptr);
        for (int i = 0; i < 2000; i += 9) { mA[i] = &A.held2[i]; mB[i] = &B.held2[i]; }
        mA[5] = A.vp[17]; mB[5] = B.vp[17]; // a point of the list that is already matched
        ORBmatcher ref(0.9f, true);
        GpuMatcher gpu(0.9f, true);
        const int nA = ref.SearchByProjection(&A.K, A.Scw, A.vp, mA, 8, 0.9f), nB = gpu.SearchByProjection(&B.K, B.Scw, B.vp, mB, 8, 0.9f);
        std::printf("SearchByProjection(KF, Sim3): ref %d gpu %d\n", nA, nB);
        EXPECT(nA == nB && tags(mA) == tags(mB) && nA > 100, "SearchByProjection(KeyFrame*, Sim3f&, ...) vpMatched / return value");
        std::vector<MapPoint *> qA(2000, nullptr), qB(2000, nullptr);
        std::vector<KeyFrame *> kA(2000, nullptr), kB(2000, nullptr);
        const int n2A = ref.SearchByProjection(&A.K, A.Scw, A.vp, A.vpKFs, qA, kA, 6, 1.0f), n2B = gpu.SearchByProjection(&B.K, B.Scw, B.vp, B.vpKFs, qB, kB, 6, 1.0f);
        bool kfs_same = true;
        for (int i = 0; i < 2000; i++) kfs_same = kfs_same && ((kA[i] == nullptr) == (kB[i] == nullptr)) && ((kA[i] == &A.K) == (kB[i] == &B.K));
        std::printf("SearchByProjection(KF, Sim3, KFs): ref %d gpu %d\n", n2A, n2B);
        EXPECT(n2A == n2B && tags(qA) == tags(qB) && kfs_same && n2A > 100, "SearchByProjection(KeyFrame*, Sim3f&, ..., KFs) vpMatched / vpMatchedKF / return value");
    }
    for (int stereo = 0; stereo < 2; stereo++)
    {
        Scene A, B;
        build_scene(A, 83 + stereo, stereo != 0); build_scene(B, 83 + stereo, stereo != 0);
        for (Scene *S : {&A, &B}) { S->vp[11] = nullptr; S->vp[40] = S->vp[39]; } // a NULL entry and a duplicate
        ORBmatcher ref(0.9f, true);
        GpuMatcher gpu(0.9f, true);
        g_replace_log().clear();
        const int nA = ref.Fuse(&A.K, A.vp, 3.0f, false);
        std::vector<std::pair<int, int>> logA, logB;
        for (auto &e : g_replace_log()) logA.push_back(std::make_pair(tag_of(e.first), tag_of(e.second)));
        g_replace_log().clear();
        const int nB = gpu.Fuse(&B.K, B.vp, 3.0f, false);
        for (auto &e : g_replace_log()) logB.push_back(std::make_pair(tag_of(e.first), tag_of(e.second)));
        std::printf("Fuse stereo=%d: ref %d gpu %d (%zu replacements)\n", stereo, nA, nB, logA.size());
        EXPECT(nA == nB && tags(A.K.mvpMapPoints) == tags(B.K.mvpMapPoints) && logA == logB && nA > 100, "Fuse(KeyFrame*, vector<MapPoint*>) map-point table / Replace calls / return value");
    }
    {
        Scene A, B;
        build_scene(A, 85, false); build_scene(B, 85, false);
        std::vector<MapPoint *> rA(A.vp.size(), nullptr), rB(B.vp.size(), nullptr);
        ORBmatcher ref(0.9f, true);
        GpuMatcher gpu(0.9f, true);
        const int nA = ref.Fuse(&A.K, A.Scw, A.vp, 4.0f, rA), nB = gpu.Fuse(&B.K, B.Scw, B.vp, 4.0f, rB);
        std::printf("Fuse(Sim3): ref %d gpu %d\n", nA, nB);
        EXPECT(nA == nB && tags(rA) == tags(rB) && tags(A.K.mvpMapPoints) == tags(B.K.mvpMapPoints) && nA > 100, "Fuse(KeyFrame*, Sim3f&, ...) vpReplacePoint / map-point table / return value");
    }
    {
        Scene A, B;
        build_scene(A, 86, false); build_scene(B, 86, false);
        for (Scene *S : {&A, &B})
        {   // K holds each landmark at its keypoint; K2 holds a twin map point (same position and descriptor) at the keypoint
            // the landmark projects to there, so that the two directions can agree; the rest is clutter
            for (int i = 0; i < 2000; i++) { S->K.mvpMapPoints[i] = nullptr; S->K2.mvpMapPoints[i] = nullptr; }
            S->twins.resize(S->pair_m.size());
            for (size_t k = 0; k < S->pair_m.size(); k++)
            {
                S->twins[k] = S->pts[S->pair_m[k]];
                S->twins[k].tag = 300000 + S->pair_m[k];
                S->twins[k].bad_ = false;
                S->K.mvpMapPoints[S->pair_tgt[k]] = S->vp[S->pair_m[k]];
                S->K2.mvpMapPoints[S->pair_j[k]] = &S->twins[k];
            }
            for (int m = 0; m < 600; m++)
            {
                const int i = (m * 37) % 2000;
                if (!S->K.mvpMapPoints[i]) S->K.mvpMapPoints[i] = S->vp[(m * 5 + 1) % (int)S->vp.size()];
            }
        }
        std::vector<MapPoint *> mA(2000, nullptr), mB(2000, nullptr);
        ORBmatcher ref(0.9f, true);
        GpuMatcher gpu(0.9f, true);
        const int nA = ref.SearchBySim3(&A.K, &A.K2, mA, A.S12, 7.5f), nB = gpu.SearchBySim3(&B.K, &B.K2, mB, B.S12, 7.5f);
        std::printf("SearchBySim3: ref %d gpu %d\n", nA, nB);
        EXPECT(nA == nB && tags(mA) == tags(mB) && nA > 50, "SearchBySim3 vpMatches12 / return value");
    }
    {
        cv::Mat a(1, 32, CV_8U), b(1, 32, CV_8U);
        for (int i = 0; i < 32; i++) { a.ptr<uint8_t>()[i] = (uint8_t)(rng() & 0xFF); b.ptr<uint8_t>()[i] = (uint8_t)(rng() & 0xFF); }
        EXPECT(ORBmatcher::DescriptorDistance(a, b) == GpuMatcher::DescriptorDistance(a, b), "DescriptorDistance");
    }
    // ---- Tracking::SearchLocalPoints: Frame::isInFrustum (the reference's own text) + SearchByProjection on the host, against the
    // adapter's single device call, with the adapter's checker counting every isInFrustum disagreement
    {
        Scene A, B;
        build_scene(A, 91, true); build_scene(B, 91, true);
        int toMatchA = 0, toMatchB = 0, nA = 0, nB = 0;
        for (Scene *S : {&A, &B})
        {
            Frame &F = S->Cur;
            F.SetPose(F.mTcw); // fills mRcw / mtcw / mOw like Frame::UpdatePoseMatrices
            for (int i = 0; i < F.N; i++) F.mvpMapPoints[i] = nullptr;
            for (size_t i = 0; i < S->vp.size(); i++)
            {
                MapPoint *p = S->vp[i];
                p->mnLastFrameSeen = (i % 11 == 3) ? F.mnId : 0; // already matched in this frame: not tested (Tracking.cc:4133)
                p->mbTrackInView = false;
                const float d = (p->worldPos_ - F.mOw).norm();
                p->mfMaxDistance = d * F.mvScaleFactors[i % 8] * (i % 13 == 5 ? 0.4f : 1.0f); // some beyond the distance invariance
                p->mfMinDistance = p->mfMaxDistance / 3.6f;
                Eigen::Vector3f nrm = (p->worldPos_ - F.mOw) / d;
                if (i % 9 == 4) nrm = Eigen::Vector3f(nrm(1), -nrm(0), 0.2f); // outside the viewing-angle gate
                p->normal_ = nrm;
            }
        }
        {
            Frame &F = A.Cur; // the reference's loop (Tracking.cc:4125-4182)
            for (MapPoint *p : A.vp)
            {
                if (p->mnLastFrameSeen == F.mnId) continue;
                if (p->isBad()) continue;
                if (F.isInFrustum(p, 0.5f)) toMatchA++;
            }
            ORBmatcher ref(0.8f, true);
            nA = ref.SearchByProjection(F, A.vp, 3.0f, false, 50.0f);
        }
        {
            orbgpu::frustum_report() = orbgpu::FrustumReport();
            orbgpu::frustum_report().enabled = true;
            GpuMatcher gpu(0.8f, true);
            nB = gpu.SearchLocalPoints(B.Cur, B.vp, 3.0f, false, 50.0f, 0.5f, &toMatchB);
            orbgpu::frustum_report().enabled = false;
        }
        bool iv = true;
        for (size_t i = 0; i < A.vp.size(); i++) iv = iv && A.vp[i]->mbTrackInView == B.vp[i]->mbTrackInView;
        const orbgpu::FrustumReport &rep = orbgpu::frustum_report();
        std::printf("SearchLocalPoints: ref %d matches / %d in view, gpu %d / %d; isInFrustum checker: %lu points, %lu in-view, %lu level, %lu projection disagreements\n",
                    nA, toMatchA, nB, toMatchB, rep.points, rep.in_view_mismatch, rep.level_mismatch, rep.projection_mismatch);
        EXPECT(nA == nB && toMatchA == toMatchB && iv && tags(A.Cur.mvpMapPoints) == tags(B.Cur.mvpMapPoints) && nA > 100 && toMatchA > 300,
               "SearchLocalPoints (isInFrustum + SearchByProjection fused) F.mvpMapPoints / mbTrackInView / counts");
        EXPECT(rep.points > 1000 && rep.in_view_mismatch == 0 && rep.level_mismatch == 0 && rep.projection_mismatch == 0,
               "device isInFrustum vs the reference's host isInFrustum: zero disagreements");
    }
    // ---- ORBVocabulary::transform (TemplatedVocabulary.h:1127-1194): the reference's own code against the drop-in override
    {
        typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVoc;
        typedef orbgpu::ORBVocabularyT<DBoW2::FORB::TDescriptor, DBoW2::FORB> GpuVoc;
        DUtils::Random::SeedRandOnce(7);
        // training set: 40 "images" of 300 descriptors drawn around 4000 prototypes (diverse enough that the reference's k-means
        // never meets an empty cluster, which it does not guard against)
        std::vector<cv::Mat> proto(4000);
        for (auto &m : proto) { m.create(1, 32, CV_8U); for (int b = 0; b < 32; b++) m.ptr<uint8_t>()[b] = (uint8_t)(rng() & 0xFF); }
        auto noisy = [&](int flips) {
            cv::Mat d = proto[rng() % proto.size()].clone();
            for (int f = 0; f < flips; f++) { int bit = (int)(rng() % 256); d.ptr<uint8_t>()[bit >> 3] ^= (uint8_t)(1 << (bit & 7)); }
            return d;
        };
        std::vector<std::vector<cv::Mat>> training(40);
        for (auto &img : training) for (int i = 0; i < 300; i++) img.push_back(noisy(8 + (int)(rng() % 30)));
        GpuVoc voc(6, 3, DBoW2::TF_IDF, DBoW2::L1_NORM);
        voc.create(training); // the reference's create() (inherited)
        std::vector<cv::Mat> feats;
        for (int i = 0; i < 2000; i++) feats.push_back(noisy((int)(rng() % 20)));
        bool all = true;
        size_t words = 0, nodes = 0;
        for (int levelsup = 0; levelsup <= 5; levelsup++)
        {
            DBoW2::BowVector vA, vB;
            DBoW2::FeatureVector fA, fB;
            voc.RefVoc::transform(feats, vA, fA, levelsup); // the reference's implementation, statically bound
            voc.transform(feats, vB, fB, levelsup);         // the device override (virtual dispatch)
            all = all && vA == vB && fA == fB && !vA.empty();
            words = vA.size(); nodes = fA.size();
        }
        std::printf("ORBVocabulary::transform: %u-word vocabulary, 2000 features -> %zu words, %zu nodes at levelsup 5; bit-exact at levelsup 0..5: %d\n",
                    voc.size(), words, nodes, (int)all);
        EXPECT(all && voc.size() > 100, "ORBVocabulary::transform BowVector (bit-exact doubles) / FeatureVector maps");
        DBoW2::BowVector e1; DBoW2::FeatureVector e2;
        voc.transform(std::vector<cv::Mat>(), e1, e2, 4);
        EXPECT(e1.empty() && e2.empty(), "transform of an empty feature list");
    }
    // the thread's device-frame cache served the repeated key frames / frames of the cases above
    std::printf("device frame cache: %lu hits, %lu misses\n", orbgpu::frame_cache().hits, orbgpu::frame_cache().misses);
    EXPECT(orbgpu::frame_cache().hits > 0, "frame cache hits");
    { // SearchByNN (defined by this repo): against a loop over the reference's own DescriptorDistance with SearchByBoW's accept rule
        const int nq = 700, nd = 3000;
        cv::Mat Q(nq, 32, CV_8U), D(nd, 32, CV_8U);
        for (int i = 0; i < nd * 32; i++) D.ptr<uint8_t>()[i] = (uint8_t)(rng() & 0xFF);
        for (int i = 0; i < nq * 32; i++) Q.ptr<uint8_t>()[i] = (uint8_t)(rng() & 0xFF);
        for (int q = 0; q < nq; q += 2) { // planted: a database row with a few flipped bits; every 10th has an exact twin (ratio failure)
            const int src = (q * 13) % nd;
            std::memcpy(Q.ptr<uint8_t>(q), D.ptr<uint8_t>(src), 32);
            for (int f = 0; f < (q % 7); f++) Q.ptr<uint8_t>(q)[(q + 5 * f) % 32] ^= (uint8_t)(1u << (f % 8));
            if (q % 10 == 0) std::memcpy(D.ptr<uint8_t>((src + 1) % nd), D.ptr<uint8_t>(src), 32);
        }
        const float ratio = 0.8f;
        std::vector<int> exp(nq, -1), got;
        int nExp = 0;
        for (int q = 0; q < nq; q++) {
            int best = 256, second = 256, bi = -1;
            for (int d = 0; d < nd; d++) {
                const int dist = ORBmatcher::DescriptorDistance(Q.row(q), D.row(d));
                if (dist < best) { second = best; best = dist; bi = d; }
                else if (dist < second) second = dist;
            }
            if (best <= ORBmatcher::TH_LOW && (float)best < ratio * (float)second) { exp[q] = bi; nExp++; }
        }
        GpuMatcher gpu(ratio, true);
        const int nGot = gpu.SearchByNN(Q, D, got);
        std::printf("SearchByNN: ref %d gpu %d\n", nExp, nGot);
        EXPECT(nGot == nExp && got == exp && nExp > 100, "SearchByNN vnMatches / return value");
        // a second call with a smaller, different database goes through the thread's resident database object (chunked re-upload)
        const int nd2 = 2000;
        cv::Mat D2(nd2, 32, CV_8U);
        for (int d = 0; d < nd2; d++) std::memcpy(D2.ptr<uint8_t>(d), D.ptr<uint8_t>(nd - 1 - d), 32);
        std::vector<int> exp2(nq, -1), got2;
        int nExp2 = 0;
        for (int q = 0; q < nq; q++) {
            int best = 256, second = 256, bi = -1;
            for (int d = 0; d < nd2; d++) {
                const int dist = ORBmatcher::DescriptorDistance(Q.row(q), D2.row(d));
                if (dist < best) { second = best; best = dist; bi = d; }
                else if (dist < second) second = dist;
            }
            if (best <= ORBmatcher::TH_LOW && (float)best < ratio * (float)second) { exp2[q] = bi; nExp2++; }
        }
        const int nGot2 = gpu.SearchByNN(Q, D2, got2);
        std::printf("SearchByNN (second database, re-upload): ref %d gpu %d\n", nExp2, nGot2);
        EXPECT(nGot2 == nExp2 && got2 == exp2 && nExp2 > 50, "SearchByNN on a re-uploaded database");
    }
    std::printf(fails ? "ADAPTER PARITY FAILED (%d)\n" : "ADAPTER PARITY OK (%d failures)\n", fails);
    return fails ? 1 : 0;
}
