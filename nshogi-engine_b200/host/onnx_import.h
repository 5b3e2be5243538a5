// onnx_import.h — reads the reference's model files for infer::B200::load, without protobuf or ONNX libraries.
//
// The reference hands an ONNX file to TensorRT's parser and builds an engine (reference src/infer/trt.cc:109-183);
// the tensor contract of that file is pinned there: input `input` [B,86,9,9] (:144-150), outputs `policy` (2187
// values per sample, :193-214), `value`, `draw` (:215-227).  This executor has no engine build step: the file's
// weights go straight into the canonical fp32 blob.  A small protobuf wire-format decoder reads
// the ModelProto; a walk over the graph recognises the ResNet the kernels run - stem conv3x3, residual blocks
// (conv-[bn]-relu-conv-[bn]-add-relu), policy head conv1x1(27), value head conv1x1(1)-[bn]-relu-fc-relu-fc-sigmoid
// -> value / draw - and folds batch-norm in double precision.  Anything else in the graph is an error: a net the
// executor cannot run must not load.  C++ twin of nshogi-engine_b200/onnx_io.py (same checks, bit-identical blob;
// tests/test_onnx_io.py compares the two on the file written by torch.onnx.export).
#ifndef NSHOGI_ENGINE_INFER_ONNX_IMPORT_H
#define NSHOGI_ENGINE_INFER_ONNX_IMPORT_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace nshogi {
namespace engine {
namespace infer {
namespace onnx {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- protobuf wire format -----------------------------------------------------------------------
struct Span {
    const uint8_t* P = nullptr;
    std::size_t N = 0;
};

struct Field {
    uint32_t No = 0;
    int Wire = 0;
    uint64_t Value = 0;  // varint / fixed
    Span Bytes;          // length-delimited
};

class Reader {
 public:
    explicit Reader(Span S) : Cur(S.P), End(S.P + S.N) {
    }
    bool next(Field* F) {
        if (Cur >= End) return false;
        const uint64_t Key = varint();
        F->No = (uint32_t)(Key >> 3);
        F->Wire = (int)(Key & 7);
        F->Bytes = Span{};
        switch (F->Wire) {
        case 0: F->Value = varint(); break;
        case 1: F->Value = fixed(8); break;
        case 5: F->Value = fixed(4); break;
        case 2: {
            const uint64_t Len = varint();
            if (Len > (uint64_t)(End - Cur)) throw Error("onnx: truncated protobuf message");
            F->Bytes = Span{Cur, (std::size_t)Len};
            Cur += Len;
            break;
        }
        default: throw Error("onnx: unsupported protobuf wire type");
        }
        return true;
    }
    uint64_t varint() {
        uint64_t V = 0;
        for (int Shift = 0; Shift < 70; Shift += 7) {
            if (Cur >= End) throw Error("onnx: truncated varint");
            const uint8_t B = *Cur++;
            V |= (uint64_t)(B & 0x7F) << Shift;
            if (B < 0x80) return V;
        }
        throw Error("onnx: malformed varint");
    }
    bool done() const {
        return Cur >= End;
    }

 private:
    uint64_t fixed(int Bytes) {
        if (End - Cur < Bytes) throw Error("onnx: truncated fixed-width field");
        uint64_t V = 0;
        std::memcpy(&V, Cur, (std::size_t)Bytes);  // little-endian hosts only (x86-64 / aarch64)
        Cur += Bytes;
        return V;
    }
    const uint8_t* Cur;
    const uint8_t* End;
};

inline std::string str(Span S) {
    return std::string(reinterpret_cast<const char*>(S.P), S.N);
}

inline void repeatedInt64(std::vector<int64_t>* Out, const Field& F) {
    if (F.Wire == 0) {
        Out->push_back((int64_t)F.Value);
    } else {
        Reader R(F.Bytes);
        while (!R.done()) Out->push_back((int64_t)R.varint());
    }
}

// ---- the subset of ONNX these graphs use ----------------------------------------------------------
struct Tensor {
    std::string Name;
    std::vector<int64_t> Dims;
    std::vector<double> Data;  // float / double / float16 / int32 / int64 payloads, widened
    bool Present = false;
    std::size_t size() const {
        return Data.size();
    }
};

struct Attr {
    bool HasF = false, HasI = false, HasS = false, HasT = false;
    float F = 0.f;
    int64_t I = 0;
    std::string S;
    Tensor T;
    std::vector<int64_t> Ints;
};

struct Node {
    std::string Op, Name;
    std::vector<std::string> In, Out;
    std::map<std::string, Attr> Attrs;
    bool Used = false;
    int64_t attrInt(const std::string& K, int64_t Default) const {
        auto It = Attrs.find(K);
        return It != Attrs.end() && It->second.HasI ? It->second.I : Default;
    }
    std::vector<int64_t> attrInts(const std::string& K, std::vector<int64_t> Default) const {
        auto It = Attrs.find(K);
        return It != Attrs.end() ? It->second.Ints : Default;
    }
};

struct Graph {
    std::vector<Node> Nodes;
    std::map<std::string, Tensor> Constants;  // initializers + Constant nodes
    std::vector<std::string> Inputs, Outputs;
};

inline float halfToFloat(uint16_t H) {
    const uint32_t Sign = (uint32_t)(H >> 15) << 31, Exp = (H >> 10) & 31, Man = H & 1023;
    uint32_t Bits;
    if (Exp == 0) {
        if (Man == 0) {
            Bits = Sign;
        } else {
            int E = -1;
            uint32_t M = Man;
            do { M <<= 1; ++E; } while (!(M & 1024));
            Bits = Sign | ((uint32_t)(127 - 15 - E) << 23) | ((M & 1023) << 13);
        }
    } else if (Exp == 31) {
        Bits = Sign | 0x7F800000u | (Man << 13);
    } else {
        Bits = Sign | ((Exp + 112) << 23) | (Man << 13);
    }
    float F;
    std::memcpy(&F, &Bits, 4);
    return F;
}

inline Tensor parseTensor(Span S) {
    Tensor T;
    T.Present = true;
    int DType = 1;
    Span Raw;
    bool HasRaw = false;
    std::vector<double> Floats, Doubles;
    std::vector<int64_t> I32, I64;
    Reader R(S);
    Field F;
    while (R.next(&F)) {
        switch (F.No) {
        case 1: repeatedInt64(&T.Dims, F); break;
        case 2: DType = (int)F.Value; break;
        case 8: T.Name = str(F.Bytes); break;
        case 9: Raw = F.Bytes; HasRaw = true; break;
        case 4:
            if (F.Wire == 2) {
                for (std::size_t I = 0; I + 4 <= F.Bytes.N; I += 4) {
                    float V;
                    std::memcpy(&V, F.Bytes.P + I, 4);
                    Floats.push_back(V);
                }
            } else {
                float V;
                const uint32_t B = (uint32_t)F.Value;
                std::memcpy(&V, &B, 4);
                Floats.push_back(V);
            }
            break;
        case 10:
            if (F.Wire == 2) {
                for (std::size_t I = 0; I + 8 <= F.Bytes.N; I += 8) {
                    double V;
                    std::memcpy(&V, F.Bytes.P + I, 8);
                    Doubles.push_back(V);
                }
            } else {
                double V;
                std::memcpy(&V, &F.Value, 8);
                Doubles.push_back(V);
            }
            break;
        case 5: repeatedInt64(&I32, F); break;
        case 7: repeatedInt64(&I64, F); break;
        case 14:
            if (F.Value == 1) throw Error("onnx: tensor '" + T.Name + "' uses external data (export with weights embedded)");
            break;
        default: break;
        }
    }
    auto fromRaw = [&](std::size_t Width, auto Convert) {
        if (Raw.N % Width) throw Error("onnx: tensor '" + T.Name + "': raw data size");
        for (std::size_t I = 0; I < Raw.N; I += Width) T.Data.push_back(Convert(Raw.P + I));
    };
    switch (DType) {
    case 1:
        if (HasRaw) fromRaw(4, [](const uint8_t* P) { float V; std::memcpy(&V, P, 4); return (double)V; });
        else T.Data = Floats;
        break;
    case 11:
        if (HasRaw) fromRaw(8, [](const uint8_t* P) { double V; std::memcpy(&V, P, 8); return V; });
        else T.Data = Doubles;
        break;
    case 7:
        if (HasRaw) fromRaw(8, [](const uint8_t* P) { int64_t V; std::memcpy(&V, P, 8); return (double)V; });
        else T.Data.assign(I64.begin(), I64.end());
        break;
    case 6:
        if (HasRaw) fromRaw(4, [](const uint8_t* P) { int32_t V; std::memcpy(&V, P, 4); return (double)V; });
        else T.Data.assign(I32.begin(), I32.end());
        break;
    case 10:
        if (HasRaw) fromRaw(2, [](const uint8_t* P) { uint16_t V; std::memcpy(&V, P, 2); return (double)halfToFloat(V); });
        else for (int64_t V : I32) T.Data.push_back((double)halfToFloat((uint16_t)V));
        break;
    default: throw Error("onnx: tensor '" + T.Name + "': unsupported data type " + std::to_string(DType));
    }
    // an initializer of this model family has small positive dims (a scalar: none); anything else - zero, negative, or a
    // product that would wrap - is a damaged file, not a shape to reinterpret
    std::size_t Count = 1;
    for (int64_t D : T.Dims) {
        if (D <= 0 || D > (int64_t(1) << 24) || Count > (std::size_t(1) << 40) / (std::size_t)D)
            throw Error("onnx: tensor '" + T.Name + "': dimension " + std::to_string(D) + " out of range");
        Count *= (std::size_t)D;
    }
    if (T.Data.size() != Count) throw Error("onnx: tensor '" + T.Name + "': value count does not match its shape");
    return T;
}

inline Node parseNode(Span S) {
    Node N;
    Reader R(S);
    Field F;
    while (R.next(&F)) {
        switch (F.No) {
        case 1: N.In.push_back(str(F.Bytes)); break;
        case 2: N.Out.push_back(str(F.Bytes)); break;
        case 3: N.Name = str(F.Bytes); break;
        case 4: N.Op = str(F.Bytes); break;
        case 5: {
            Attr A;
            std::string Key;
            Reader RA(F.Bytes);
            Field FA;
            while (RA.next(&FA)) {
                switch (FA.No) {
                case 1: Key = str(FA.Bytes); break;
                case 2: { const uint32_t B = (uint32_t)FA.Value; std::memcpy(&A.F, &B, 4); A.HasF = true; break; }
                case 3: A.I = (int64_t)FA.Value; A.HasI = true; break;
                case 4: A.S = str(FA.Bytes); A.HasS = true; break;
                case 5: A.T = parseTensor(FA.Bytes); A.HasT = true; break;
                case 8: repeatedInt64(&A.Ints, FA); break;
                default: break;
                }
            }
            N.Attrs[Key] = std::move(A);
            break;
        }
        case 7: {
            const std::string Domain = str(F.Bytes);
            if (!Domain.empty() && Domain != "ai.onnx") throw Error("onnx: operator domain '" + Domain + "' is not supported");
            break;
        }
        default: break;
        }
    }
    return N;
}

inline std::string valueInfoName(Span S) {
    Reader R(S);
    Field F;
    while (R.next(&F))
        if (F.No == 1) return str(F.Bytes);
    return "";
}

inline Graph parseModel(const std::vector<uint8_t>& Data) {
    Span GraphBytes;
    bool HasGraph = false;
    {
        Reader R(Span{Data.data(), Data.size()});
        Field F;
        while (R.next(&F))
            if (F.No == 7 && F.Wire == 2) { GraphBytes = F.Bytes; HasGraph = true; }
    }
    if (!HasGraph) throw Error("onnx: not a ModelProto (no graph)");
    Graph G;
    std::vector<std::string> Inputs;
    Reader R(GraphBytes);
    Field F;
    while (R.next(&F)) {
        switch (F.No) {
        case 1: {
            Node N = parseNode(F.Bytes);
            if (N.Op == "Constant") {
                auto It = N.Attrs.find("value");
                if (It == N.Attrs.end() || !It->second.HasT || N.Out.empty()) throw Error("onnx: Constant node without a tensor value");
                G.Constants[N.Out[0]] = It->second.T;
            } else {
                G.Nodes.push_back(std::move(N));
            }
            break;
        }
        case 5: { Tensor T = parseTensor(F.Bytes); G.Constants[T.Name] = std::move(T); break; }
        case 11: Inputs.push_back(valueInfoName(F.Bytes)); break;
        case 12: G.Outputs.push_back(valueInfoName(F.Bytes)); break;
        default: break;
        }
    }
    for (auto& I : Inputs)
        if (!G.Constants.count(I)) G.Inputs.push_back(I);
    return G;
}

// ---- graph -> canonical blob -----------------------------------------------------------------------
struct NetShape {
    int InChannels = 0, Channels = 0, Blocks = 0, Hidden = 0;
};

class Walker {
 public:
    explicit Walker(Graph& Gr) : G(Gr) {
        for (auto& N : G.Nodes)
            for (auto& X : N.In)
                if (!X.empty() && !G.Constants.count(X)) Consumers[X].push_back(&N);
    }

    std::vector<Node*> cons(const std::string& Name, const char* Op = nullptr) {
        std::vector<Node*> R;
        auto It = Consumers.find(Name);
        if (It == Consumers.end()) return R;
        for (Node* N : It->second)
            if (!Op || N->Op == Op) R.push_back(N);
        return R;
    }
    Node* sole(const std::string& Name, const char* Op, const std::string& What) {
        auto C = cons(Name);
        if (C.size() != 1 || C[0]->Op != Op) throw Error("onnx: " + What + ": expected a single " + Op + " after '" + Name + "'");
        C[0]->Used = true;
        return C[0];
    }
    const Tensor& weight(const std::string& Name, const std::string& What) {
        auto It = G.Constants.find(Name);
        if (It == G.Constants.end()) throw Error("onnx: " + What + ": '" + Name + "' is not a constant tensor");
        return It->second;
    }

    // Conv [+ BatchNormalization] -> folded weight [Cout][Cin][k][k] and bias, returns the output tensor name
    std::string conv(Node* N, int K, const std::string& What, std::vector<double>* W, std::vector<double>* B, int* Cout,
                     int* Cin) {
        N->Used = true;
        const Tensor& Wt = weight(N->In.at(1), What);
        if (Wt.Dims.size() != 4 || Wt.Dims[2] != K || Wt.Dims[3] != K)
            throw Error("onnx: " + What + ": expected a " + std::to_string(K) + "x" + std::to_string(K) + " convolution");
        const int64_t Pad = (K - 1) / 2;
        const auto Pads = N->attrInts("pads", {0, 0, 0, 0});
        std::string AutoPad = "NOTSET";
        if (auto It = N->Attrs.find("auto_pad"); It != N->Attrs.end() && It->second.HasS) AutoPad = It->second.S;
        const bool PadsOk = Pads == std::vector<int64_t>{Pad, Pad, Pad, Pad} || AutoPad == "SAME_UPPER" || AutoPad == "SAME_LOWER" ||
                            (Pad == 0 && AutoPad == "VALID");
        if (!PadsOk) throw Error("onnx: " + What + ": needs 'same' padding");
        if (N->attrInts("strides", {1, 1}) != std::vector<int64_t>{1, 1} || N->attrInts("dilations", {1, 1}) != std::vector<int64_t>{1, 1} ||
            N->attrInt("group", 1) != 1)
            throw Error("onnx: " + What + ": stride / dilation / group must be 1");
        *Cout = (int)Wt.Dims[0];
        *Cin = (int)Wt.Dims[1];
        *W = Wt.Data;
        B->assign((std::size_t)*Cout, 0.0);
        if (N->In.size() > 2 && !N->In[2].empty()) {
            const Tensor& Bt = weight(N->In[2], What);
            if (Bt.size() != (std::size_t)*Cout) throw Error("onnx: " + What + ": bias size");
            *B = Bt.Data;
        }
        std::string Out = N->Out.at(0);
        auto Next = cons(Out);
        if (Next.size() == 1 && Next[0]->Op == "BatchNormalization") {
            Node* Bn = Next[0];
            Bn->Used = true;
            if (Bn->attrInt("training_mode", 0)) throw Error("onnx: " + What + ": batch-norm in training mode; export in eval mode");
            float EpsF = 1e-5f;
            if (auto It = Bn->Attrs.find("epsilon"); It != Bn->Attrs.end() && It->second.HasF) EpsF = It->second.F;
            const double Eps = (double)EpsF;
            const Tensor &Gamma = weight(Bn->In.at(1), What), &Beta = weight(Bn->In.at(2), What), &Mean = weight(Bn->In.at(3), What),
                         &Var = weight(Bn->In.at(4), What);
            const std::size_t Per = W->size() / (std::size_t)*Cout;
            for (int C = 0; C < *Cout; ++C) {  // weights_io.fold_bn
                const double Scale = Gamma.Data.at(C) / std::sqrt(Var.Data.at(C) + Eps);
                for (std::size_t I = 0; I < Per; ++I) (*W)[C * Per + I] *= Scale;
                (*B)[C] = ((*B)[C] - Mean.Data.at(C)) * Scale + Beta.Data.at(C);
            }
            Out = Bn->Out.at(0);
        }
        return Out;
    }
    std::string relu(const std::string& Name, const std::string& What) {
        return sole(Name, "Relu", What)->Out.at(0);
    }
    std::string through(std::string Name) {
        static const std::set<std::string> Pass = {"Flatten", "Reshape", "Identity", "Squeeze", "Unsqueeze"};
        for (;;) {
            auto C = cons(Name);
            if (C.size() == 1 && Pass.count(C[0]->Op)) {
                C[0]->Used = true;
                Name = C[0]->Out.at(0);
                continue;
            }
            return Name;
        }
    }
    // Gemm, or MatMul + Add -> weight [Out][In], bias [Out]
    std::string dense(Node* N, const std::string& What, std::vector<double>* W, std::vector<double>* B, int* Out, int* In) {
        N->Used = true;
        if (N->Op == "Gemm") {
            float Alpha = 1.f, Beta = 1.f;
            if (auto It = N->Attrs.find("alpha"); It != N->Attrs.end() && It->second.HasF) Alpha = It->second.F;
            if (auto It = N->Attrs.find("beta"); It != N->Attrs.end() && It->second.HasF) Beta = It->second.F;
            if (N->attrInt("transA", 0) || Alpha != 1.f || Beta != 1.f) throw Error("onnx: " + What + ": Gemm with transA / alpha / beta");
            const Tensor& Wt = weight(N->In.at(1), What);
            if (Wt.Dims.size() != 2) throw Error("onnx: " + What + ": Gemm weight rank");
            const bool TransB = N->attrInt("transB", 0) != 0;
            *Out = (int)(TransB ? Wt.Dims[0] : Wt.Dims[1]);
            *In = (int)(TransB ? Wt.Dims[1] : Wt.Dims[0]);
            W->resize(Wt.size());
            for (int O = 0; O < *Out; ++O)
                for (int I = 0; I < *In; ++I) (*W)[(std::size_t)O * *In + I] = TransB ? Wt.Data[(std::size_t)O * *In + I] : Wt.Data[(std::size_t)I * *Out + O];
            B->assign((std::size_t)*Out, 0.0);
            if (N->In.size() > 2 && !N->In[2].empty()) *B = weight(N->In[2], What).Data;
            if (B->size() != (std::size_t)*Out) throw Error("onnx: " + What + ": bias size");
            return N->Out.at(0);
        }
        if (N->Op == "MatMul") {
            const Tensor& Wt = weight(N->In.at(1), What);
            if (Wt.Dims.size() != 2) throw Error("onnx: " + What + ": MatMul weight rank");
            *In = (int)Wt.Dims[0];
            *Out = (int)Wt.Dims[1];
            W->resize(Wt.size());
            for (int O = 0; O < *Out; ++O)
                for (int I = 0; I < *In; ++I) (*W)[(std::size_t)O * *In + I] = Wt.Data[(std::size_t)I * *Out + O];
            B->assign((std::size_t)*Out, 0.0);
            std::string Res = N->Out.at(0);
            auto Add = cons(Res);
            if (Add.size() == 1 && Add[0]->Op == "Add") {
                std::vector<std::string> Other;
                for (auto& X : Add[0]->In)
                    if (X != Res) Other.push_back(X);
                if (Other.size() == 1 && G.Constants.count(Other[0])) {
                    Add[0]->Used = true;
                    *B = weight(Other[0], What).Data;
                    if (B->size() != (std::size_t)*Out) throw Error("onnx: " + What + ": bias size");
                    Res = Add[0]->Out.at(0);
                }
            }
            return Res;
        }
        throw Error("onnx: " + What + ": expected Gemm or MatMul, found " + N->Op);
    }
    int64_t intConst(const std::string& Name, const std::string& What) {
        const Tensor& T = weight(Name, What);
        if (T.size() != 1) throw Error("onnx: " + What + ": expected a scalar index");
        return (int64_t)T.Data[0];
    }

    Graph& G;
    std::map<std::string, std::vector<Node*>> Consumers;
};

inline void append(std::vector<float>* Blob, const std::vector<double>& V) {
    for (double X : V) Blob->push_back((float)X);
}

// Canonical blob (layout: DESIGN.md §5 = weights_io.blob_from_state) of the ResNet in `G`.
inline NetShape toBlob(Graph& G, std::vector<float>* Blob) {
    auto has = [](const std::vector<std::string>& V, const char* S) {
        for (auto& X : V)
            if (X == S) return true;
        return false;
    };
    if (!has(G.Inputs, "input")) throw Error("onnx: graph input 'input' not found (reference src/infer/trt.cc:144-150)");
    for (const char* O : {"policy", "value", "draw"})
        if (!has(G.Outputs, O)) throw Error(std::string("onnx: graph output '") + O + "' not found (reference src/infer/trt.cc:193-227)");
    Walker Wk(G);
    Blob->clear();
    NetShape S;
    std::vector<double> W, B;
    int Cout = 0, Cin = 0;

    auto Stem = Wk.cons("input", "Conv");
    if (Stem.size() != 1 || Wk.cons("input").size() != 1) throw Error("onnx: the stem must be one 3x3 convolution on 'input'");
    std::string X = Wk.relu(Wk.conv(Stem[0], 3, "stem", &W, &B, &Cout, &Cin), "stem");
    S.Channels = Cout;
    S.InChannels = Cin;
    append(Blob, W);
    append(Blob, B);

    while (!Wk.cons(X, "Add").empty()) {
        const std::string What = "block " + std::to_string(S.Blocks);
        auto Add = Wk.cons(X, "Add");
        auto C1 = Wk.cons(X, "Conv");
        if (Add.size() != 1 || C1.size() != 1 || Wk.cons(X).size() != 2)
            throw Error("onnx: " + What + ": a residual block input feeds exactly one Conv and one Add");
        std::string H = Wk.relu(Wk.conv(C1[0], 3, What + " conv1", &W, &B, &Cout, &Cin), What + " conv1");
        if (Cout != S.Channels || Cin != S.Channels) throw Error("onnx: " + What + ": channel count changes inside the trunk");
        append(Blob, W);
        append(Blob, B);
        Node* C2 = Wk.sole(H, "Conv", What + " conv2");
        const std::string Y = Wk.conv(C2, 3, What + " conv2", &W, &B, &Cout, &Cin);
        if (Cout != S.Channels || Cin != S.Channels) throw Error("onnx: " + What + ": channel count changes inside the trunk");
        append(Blob, W);
        append(Blob, B);
        const auto& AI = Add[0]->In;
        const bool SkipOk = AI.size() == 2 && ((AI[0] == X && AI[1] == Y) || (AI[0] == Y && AI[1] == X)) && Wk.cons(Y).size() == 1;
        if (!SkipOk) throw Error("onnx: " + What + ": the skip connection must add the block input to conv2's output");
        Add[0]->Used = true;
        X = Wk.relu(Add[0]->Out.at(0), What);
        ++S.Blocks;
    }
    if (S.Blocks == 0) throw Error("onnx: no residual block found after the stem");

    auto Heads = Wk.cons(X);
    Node *PolicyConv = nullptr, *ValueConv = nullptr;
    if (Heads.size() == 2 && Heads[0]->Op == "Conv" && Heads[1]->Op == "Conv") {
        for (Node* N : Heads) {
            const Tensor& T = Wk.weight(N->In.at(1), "head");
            if (!T.Dims.empty() && T.Dims[0] == 27) PolicyConv = N;
            if (!T.Dims.empty() && T.Dims[0] == 1) ValueConv = N;
        }
    }
    if (!PolicyConv || !ValueConv)
        throw Error("onnx: the trunk output must feed the policy (27 channels) and value (1 channel) 1x1 convolutions");
    if (Wk.through(Wk.conv(PolicyConv, 1, "policy head", &W, &B, &Cout, &Cin)) != "policy")
        throw Error("onnx: policy head: conv1x1(27) must reach the output 'policy' through shape-only nodes (plane-major logits)");
    append(Blob, W);
    append(Blob, B);

    std::string V = Wk.through(Wk.relu(Wk.conv(ValueConv, 1, "value head", &W, &B, &Cout, &Cin), "value head"));
    append(Blob, W);
    append(Blob, B);
    auto Fc1 = Wk.cons(V);
    if (Fc1.size() != 1) throw Error("onnx: value head fc1: expected one fully connected layer");
    int Out = 0, In = 0;
    V = Wk.dense(Fc1[0], "value head fc1", &W, &B, &Out, &In);
    if (In != 81) throw Error("onnx: value head fc1: expected 81 inputs");
    S.Hidden = Out;
    append(Blob, W);
    append(Blob, B);
    V = Wk.relu(V, "value head fc1");

    std::vector<double> RowW[2];
    double RowB[2] = {0.0, 0.0};
    bool Have[2] = {false, false};
    auto setRow = [&](const std::string& End, const double* Wrow, double Bias) {
        const int K = End == "value" ? 0 : End == "draw" ? 1 : -1;
        if (K < 0) throw Error("onnx: value head: a component ends in '" + End + "', expected 'value' or 'draw'");
        RowW[K].assign(Wrow, Wrow + S.Hidden);
        RowB[K] = Bias;
        Have[K] = true;
    };
    auto Next = Wk.cons(V);
    if (Next.size() == 1) {
        std::string O = Wk.dense(Next[0], "value head fc2", &W, &B, &Out, &In);
        if (Out != 2 || In != S.Hidden) throw Error("onnx: value head fc2: expected weight [2, hidden]");
        O = Wk.sole(O, "Sigmoid", "value head")->Out.at(0);
        for (Node* N : Wk.cons(O)) {
            N->Used = true;
            std::vector<std::pair<int64_t, std::string>> Ends;
            const int64_t Axis = N->attrInt("axis", 0);
            if (N->Op == "Split" && (Axis == 1 || Axis == -1) && N->Out.size() == 2) {
                Ends = {{0, N->Out[0]}, {1, N->Out[1]}};
            } else if (N->Op == "Gather" && (Axis == 1 || Axis == -1)) {
                Ends = {{Wk.intConst(N->In.at(1), "value head gather"), N->Out.at(0)}};
            } else if (N->Op == "Slice") {
                const int64_t Ax = N->In.size() > 3 && !N->In[3].empty() ? Wk.intConst(N->In[3], "value head slice") : 0;
                if (Ax != 1 && Ax != -1) throw Error("onnx: value head: Slice must cut axis 1");
                Ends = {{Wk.intConst(N->In.at(1), "value head slice"), N->Out.at(0)}};
            } else {
                throw Error("onnx: value head: cannot split the (value, draw) pair with " + N->Op);
            }
            for (auto& E : Ends) {
                if (E.first != 0 && E.first != 1) throw Error("onnx: value head: component index out of range");
                setRow(Wk.through(E.second), W.data() + E.first * S.Hidden, B[(std::size_t)E.first]);
            }
        }
    } else if (Next.size() == 2) {
        for (Node* N : Next) {
            std::string O = Wk.dense(N, "value head fc2", &W, &B, &Out, &In);
            if (Out != 1 || In != S.Hidden) throw Error("onnx: value head fc2: expected weight [1, hidden]");
            setRow(Wk.through(Wk.sole(O, "Sigmoid", "value head")->Out.at(0)), W.data(), B[0]);
        }
    }
    if (!Have[0] || !Have[1]) throw Error("onnx: value head: could not resolve the outputs 'value' and 'draw'");
    append(Blob, RowW[0]);
    append(Blob, RowW[1]);
    Blob->push_back((float)RowB[0]);
    Blob->push_back((float)RowB[1]);

    for (auto& N : G.Nodes)
        if (!N.Used) throw Error("onnx: unsupported graph: node " + N.Op + "(" + (N.Name.empty() ? N.Out.at(0) : N.Name) + ") is outside the recognised ResNet");
    return S;
}

// parseModel + toBlob for a file image; whatever a damaged file trips (a missing node input, an absurd length)
// surfaces as onnx::Error like an unsupported graph does.
inline NetShape importModel(const std::vector<uint8_t>& Data, std::vector<float>* Blob) {
    try {
        Graph G = parseModel(Data);
        return toBlob(G, Blob);
    } catch (const Error&) {
        throw;
    } catch (const std::exception& E) {
        throw Error(std::string("onnx: malformed file: ") + E.what());
    }
}

} // namespace onnx
} // namespace infer
} // namespace engine
} // namespace nshogi

#endif
