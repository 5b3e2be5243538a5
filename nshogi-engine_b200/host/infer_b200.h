// infer_b200.h — `infer::B200`, the B200-native executor behind the reference's plug-in interface.
//
// Drop-in beside infer::Zero / Nothing / Random / TensorRT (reference src/infer/{zero,nothing,
// random,trt}.h): same four virtual methods (src/infer/infer.h:19-32), same construction
// pattern as TensorRT (ctor(GPUId, BatchSizeMax, NumChannels) + load() + resetGPU(),
// src/infer/trt.cc:52-80,109-232,289-291), same error behaviour (fatal: the reference exits on
// TensorRT errors, src/infer/trt.h:34-39; load failures throw std::runtime_error, trt.cc:35,130).
// Everything below the class is the C ABI of include/nsb.h (libnsb.so); there is no CPU path.
#ifndef NSHOGI_ENGINE_INFER_B200_H
#define NSHOGI_ENGINE_INFER_B200_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

#include "infer/infer.h"  // the reference's header when built in-tree, host/shim otherwise
#include "nsb.h"
#include "onnx_import.h"

namespace nshogi {
namespace engine {
namespace infer {

class B200 : public Infer {
 public:
    // NetChannels/NetBlocks select the canonical ResNet; with a weight file they
    // are read from its header instead (load()).
    B200(int GPUId, uint16_t BatchSizeMax, uint16_t NumChannels, int NetChannels = 128, int NetBlocks = 10,
         int Slots = 1)
        : BatchSizeM(BatchSizeMax), GPUId_(GPUId), Slots_(Slots) {
        static_assert(sizeof(ml::FeatureBitboard) == sizeof(nsb_feature_bitboard), "FeatureBitboard is 16 bytes");
        // 86 = preset::SimpleFeatures (globalconfig.h:19), 93 = preset::CustomFeaturesV1 (preset.h:68-122); any count the
        // stem can take is accepted, the feature set is the caller's (FeatureBitboards in, Infer contract)
        if (NumChannels < 1 || NumChannels > 96) throw std::runtime_error("B200: NumChannels must be in 1..96");
        Desc_ = nsb_net_desc{NumChannels, NetChannels, NetBlocks, 256};
        check(nsb_create(&Ctx_, GPUId, BatchSizeMax, Slots, &Desc_), "nsb_create");
    }
    ~B200() override {
        if (Ctx_) nsb_await(Ctx_, 0);
        for (void* P : Registered_) nsb_host_unregister(P);
        nsb_destroy(Ctx_);
    }
    B200(const B200&) = delete;
    B200& operator=(const B200&) = delete;

    // Loads the net.  `Path` is either the reference's ONNX model file (what TensorRT::load takes,
    // trt.cc:109-232; read by onnx_import.h, no engine build step) or an NSBW blob file ("NSBW", u32
    // version=1, i32 channels, blocks, hidden, in_channels, then the canonical fp32 blob of DESIGN.md §5;
    // written by weights_io.py).  The file decides the net's shape, as it does for the reference: if it
    // differs from the constructor's, the context is rebuilt for it.  An empty path loads the seeded
    // random-init net, which is what the benchmarks use (the reference ships no model, src/context.h:93).
    void load(const std::string& Path, uint64_t Seed = 1234) {
        std::vector<float> Blob;
        if (Path.empty()) {
            Blob.resize(nsb_weight_blob_floats(&Desc_));
            check(nsb_weight_blob_random(&Desc_, Seed, Blob.data()), "nsb_weight_blob_random");
        } else {
            std::ifstream In(Path, std::ios::binary);
            if (!In) throw std::runtime_error("B200::load: cannot open " + Path);  // trt.cc:35
            std::vector<uint8_t> Bytes((std::istreambuf_iterator<char>(In)), std::istreambuf_iterator<char>());
            nsb_net_desc FileDesc{};
            if (Bytes.size() >= 24 && std::memcmp(Bytes.data(), "NSBW", 4) == 0) {
                uint32_t Version = 0;
                int32_t H[4] = {0, 0, 0, 0};
                std::memcpy(&Version, Bytes.data() + 4, 4);
                std::memcpy(H, Bytes.data() + 8, 16);
                if (Version != 1) throw std::runtime_error("B200::load: not an NSBW v1 weight file: " + Path);
                FileDesc = nsb_net_desc{H[3], H[0], H[1], H[2]};
                const std::size_t Floats = nsb_weight_blob_floats(&FileDesc);
                if (Floats == 0 || Bytes.size() != 24 + Floats * sizeof(float))
                    throw std::runtime_error("B200::load: truncated or oversized weight file: " + Path);
                Blob.resize(Floats);
                std::memcpy(Blob.data(), Bytes.data() + 24, Floats * sizeof(float));
            } else {
                const onnx::NetShape S = onnx::importModel(Bytes, &Blob);  // throws onnx::Error (a std::runtime_error)
                FileDesc = nsb_net_desc{S.InChannels, S.Channels, S.Blocks, S.Hidden};
                if (Blob.size() != nsb_weight_blob_floats(&FileDesc))
                    throw std::runtime_error("B200::load: " + Path + ": the executor does not support this net shape");
            }
            if (FileDesc.channels != Desc_.channels || FileDesc.blocks != Desc_.blocks ||
                FileDesc.value_hidden != Desc_.value_hidden || FileDesc.in_channels != Desc_.in_channels) {
                if (HasCache_) throw std::runtime_error("B200::load: net shape changes after enableCache()");
                nsb_destroy(Ctx_);
                Ctx_ = nullptr;
                Desc_ = FileDesc;
                check(nsb_create(&Ctx_, GPUId_, BatchSizeM, Slots_, &Desc_), "nsb_create");
            }
        }
        check(nsb_load_weights(Ctx_, Blob.data(), Blob.size()), "nsb_load_weights");
    }

    // Device-resident evaluation cache shared by this executor's slots (reference: one EvalCache per
    // Manager, src/mcts/manager.cc:202-206; here it lives in HBM next to the kernels that fill it).
    void enableCache(std::size_t MemoryMiB) {
        check(nsb_cache_create(Ctx_, MemoryMiB), "nsb_cache_create");
        HasCache_ = true;
    }
    // Share another executor's cache (same GPU): the reference's evaluation workers all feed one EvalCache.
    void attachCache(B200& Owner) {
        check(nsb_cache_attach(Ctx_, Owner.Ctx_), "nsb_cache_attach");
        HasCache_ = true;
    }
    bool hasCache() const {
        return HasCache_;
    }

    void resetGPU() {  // trt.cc:289-291
        check(nsb_bind_thread(Ctx_), "nsb_bind_thread");
    }

    // The Infer contract gives the executor four caller-owned arrays that stay the same for its lifetime and are
    // sized for BatchSizeMax (src/evaluate/evaluator.cc:85-106).  The first time a set of arrays is seen it is
    // page-locked - or adopted, if the caller's Evaluator already did that - so that a one-slot executor works on
    // them directly (NSB_IO_DIRECT: no copy nodes).  Arrays that cannot be registered simply take the staged path.
    void setAutoRegister(bool On) {
        AutoRegister_ = On;
    }

    void computeNonBlocking(const ml::FeatureBitboard* Features, std::size_t BatchSize, float* DstPolicy,
                            float* DstWinRate, float* DstDrawRate) override {
        if (AutoRegister_ && (Features != Seen_[0] || DstPolicy != Seen_[1] || DstWinRate != Seen_[2] || DstDrawRate != Seen_[3])) {
            Seen_[0] = Features;
            Seen_[1] = DstPolicy;
            Seen_[2] = DstWinRate;
            Seen_[3] = DstDrawRate;
            const std::size_t B = BatchSizeM;
            const std::size_t Bytes[4] = {B * (std::size_t)Desc_.in_channels * sizeof(nsb_feature_bitboard), B * NSB_POLICY_SIZE * sizeof(float),
                                          B * sizeof(float), B * sizeof(float)};
            // Every range the library now knows - locked by this call OR adopted from the caller's Evaluator - is
            // remembered and forgotten again in ~B200 (nsb_host_unregister unlocks only what the library locked and
            // leaves nsb_host_alloc memory alone), so that no entry outlives the executor that made it.
            for (int I = 0; I < 4; ++I)
                if (nsb_host_register(const_cast<void*>(Seen_[I]), Bytes[I]) >= 0) Registered_.push_back(const_cast<void*>(Seen_[I]));
        }
        check(nsb_eval_async(Ctx_, 0, reinterpret_cast<const nsb_feature_bitboard*>(Features), BatchSize, DstPolicy,
                             DstWinRate, DstDrawRate),
              "nsb_eval_async");
    }
    void computeBlocking(const ml::FeatureBitboard* Features, std::size_t BatchSize, float* DstPolicy,
                         float* DstWinRate, float* DstDrawRate) override {
        computeNonBlocking(Features, BatchSize, DstPolicy, DstWinRate, DstDrawRate);
        await();
    }
    void await() override {
        check(nsb_await(Ctx_, 0), "nsb_await");
    }
    bool isComputing() override {
        const int R = nsb_is_computing(Ctx_, 0);
        if (R < 0) check(R, "nsb_is_computing");
        return R == 1;
    }

    // Pin the calling (evaluator) thread to the CPUs of this GPU's NUMA node: the affinity half of the reference's
    // NUMA_ENABLED Evaluator (src/evaluate/evaluator.cc:39-83); LeafPipeline places its pinned slots on that node.
    bool bindThreadToGpuNode() {
        return nsb_numa_bind_thread(GPUId_) == 0;
    }
    int gpu() const {
        return GPUId_;
    }

    nsb_ctx* context() {
        return Ctx_;
    }
    int slots() const {
        return Slots_;
    }
    const nsb_net_desc& net() const {
        return Desc_;
    }

    static void check(int Status, const char* What) {
        if (Status != NSB_OK) {  // the reference treats executor errors as fatal (trt.h:34-39)
            std::fprintf(stderr, "[infer::B200] %s failed (%d): %s\n", What, Status, nsb_last_error());
            std::exit(1);
        }
    }

 private:
    const uint16_t BatchSizeM;
    const int GPUId_;
    const int Slots_;
    nsb_net_desc Desc_;
    nsb_ctx* Ctx_ = nullptr;
    bool HasCache_ = false;
    bool AutoRegister_ = true;
    const void* Seen_[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<void*> Registered_;
};

} // namespace infer
} // namespace engine
} // namespace nshogi

#endif // NSHOGI_ENGINE_INFER_B200_H
