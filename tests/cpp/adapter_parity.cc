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

static void finish_frame(Frame &F) { F.mnMinX = 0; F.mnMinY = 0; F.mnMaxX = 640; F.mnMaxY = 480; F.mvbOutlier.assign(F.N, false); F.AssignFeaturesToGrid(0, 0); }
static void finish_kf(KeyFrame &K) { K.mnMinX = 0; K.mnMinY = 0; K.mnMaxX = 640; K.mnMaxY = 480; K.AssignFeaturesToGrid(0, 0); }

static int fails = 0;
#define EXPECT(cond, what)                                            \
    do {                                                              \
        if (!(cond)) { std::printf("MISMATCH: %s\n", what); fails++; } \
        else std::printf("ok: %s\n", what);                           \
    } while (0)

int main()
{
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
    {
        cv::Mat a(1, 32, CV_8U), b(1, 32, CV_8U);
        for (int i = 0; i < 32; i++) { a.ptr<uint8_t>()[i] = (uint8_t)(rng() & 0xFF); b.ptr<uint8_t>()[i] = (uint8_t)(rng() & 0xFF); }
        EXPECT(ORBmatcher::DescriptorDistance(a, b) == GpuMatcher::DescriptorDistance(a, b), "DescriptorDistance");
    }
    std::printf(fails ? "ADAPTER PARITY FAILED (%d)\n" : "ADAPTER PARITY OK (%d failures)\n", fails);
    return fails ? 1 : 0;
}
