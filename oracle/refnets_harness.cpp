// TEST INFRASTRUCTURE ONLY — harness around the reference's UNMODIFIED decision networks
// (/root/reference/src/neural/decisionNets.{h,cpp}), compiled against the LibTorch inside the torch
// wheel into oracle/_ref/libfastace_refnets.so.  It pins the batched re-expression of the nets
// (fastace_b200/policy.py): same parameter names, same forward results.
//
// OURS: the construction of the 11 instances mirrors DecisionNetHandler's constructor
// (src/neural/decisionNetHandler.cpp:126-222) — that class itself needs a live economy — and the
// export / forward entry points.  Everything numerical is the reference's own module code.
#include <torch/torch.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "decisionNets.h"

using namespace neural;

struct RefNets {
    int S, enc, H, nH, nHs, G;
    std::shared_ptr<OfferEncoder> offerEncoder, jobOfferEncoder;
    std::shared_ptr<PurchaseNet> purchaseNet, firmPurchaseNet, laborSearchNet;
    std::shared_ptr<ConsumptionNet> consumptionNet, productionNet;
    std::shared_ptr<OfferNet> offerNet;
    std::shared_ptr<JobOfferNet> jobOfferNet;
    std::shared_ptr<ValueNet> valueNet, firmValueNet;
    std::vector<std::pair<std::string, torch::nn::Module*>> all;
    std::vector<std::pair<std::string, torch::Tensor>> flat;   // "<net>/<param name>"
};

static torch::Tensor t1(const float* p, int64_t n) { return torch::from_blob((void*)p, {n}, torch::kFloat32).clone(); }
static torch::Tensor t2(const float* p, int64_t a, int64_t b) { return torch::from_blob((void*)p, {a, b}, torch::kFloat32).clone(); }
static void out(const torch::Tensor& t, float* dst) {
    auto c = t.contiguous().to(torch::kFloat32);
    std::memcpy(dst, c.data_ptr<float>(), sizeof(float) * c.numel());
}

extern "C" {

RefNets* refnets_create(int stackSize, int encodingSize, int hiddenSize, int nHidden, int nHiddenSmall, int numGoods,
                        uint64_t seed) {
    torch::manual_seed(seed);
    auto* r = new RefNets;
    r->S = stackSize; r->enc = encodingSize; r->H = hiddenSize; r->nH = nHidden; r->nHs = nHiddenSmall; r->G = numGoods;
    const int numUtilParams = numGoods + 3;                     // decisionNetHandler.cpp:138
    const int numProdFuncParams = numUtilParams * numGoods;     // :139
    r->offerEncoder = std::make_shared<OfferEncoder>(stackSize, numGoods + 1, hiddenSize, nHidden, encodingSize);
    r->jobOfferEncoder = std::make_shared<OfferEncoder>(stackSize, 2, hiddenSize, nHidden, encodingSize);
    r->purchaseNet = std::make_shared<PurchaseNet>(r->offerEncoder, numUtilParams, numGoods, hiddenSize, nHidden);
    r->firmPurchaseNet = std::make_shared<PurchaseNet>(r->offerEncoder, numProdFuncParams, numGoods, hiddenSize, nHidden);
    r->laborSearchNet = std::make_shared<PurchaseNet>(r->jobOfferEncoder, numUtilParams, numGoods, hiddenSize, nHidden);
    r->consumptionNet = std::make_shared<ConsumptionNet>(numUtilParams, numGoods, hiddenSize, nHidden);
    r->productionNet = std::make_shared<ConsumptionNet>(numProdFuncParams, numGoods, hiddenSize, nHidden);
    r->offerNet = std::make_shared<OfferNet>(r->offerEncoder, numProdFuncParams, numGoods, hiddenSize, hiddenSize, nHidden, nHiddenSmall);
    r->jobOfferNet = std::make_shared<JobOfferNet>(r->jobOfferEncoder, numProdFuncParams, numGoods, hiddenSize, nHidden);
    r->valueNet = std::make_shared<ValueNet>(r->offerEncoder, r->jobOfferEncoder, numUtilParams, numGoods, hiddenSize, nHidden);
    r->firmValueNet = std::make_shared<ValueNet>(r->offerEncoder, r->jobOfferEncoder, numProdFuncParams, numGoods, hiddenSize, nHidden);
    r->all = {{"offerEncoder", r->offerEncoder.get()}, {"jobOfferEncoder", r->jobOfferEncoder.get()},
              {"purchaseNet", r->purchaseNet.get()}, {"firmPurchaseNet", r->firmPurchaseNet.get()},
              {"laborSearchNet", r->laborSearchNet.get()}, {"consumptionNet", r->consumptionNet.get()},
              {"productionNet", r->productionNet.get()}, {"offerNet", r->offerNet.get()},
              {"jobOfferNet", r->jobOfferNet.get()}, {"valueNet", r->valueNet.get()},
              {"firmValueNet", r->firmValueNet.get()}};
    for (auto& kv : r->all)
        for (auto& np : kv.second->named_parameters(/*recurse=*/true))
            r->flat.push_back({kv.first + "/" + np.key(), np.value()});
    return r;
}

void refnets_destroy(RefNets* r) { delete r; }

// named_parameters() of every net, as the reference registers them (this is what torch::save writes)
int refnets_num_params(RefNets* r) { return (int)r->flat.size(); }
int refnets_param_info(RefNets* r, int i, char* name, int name_cap, int64_t* shape2) {
    std::snprintf(name, name_cap, "%s", r->flat[i].first.c_str());
    const auto& t = r->flat[i].second;
    shape2[0] = t.size(0);
    shape2[1] = t.dim() > 1 ? t.size(1) : 0;
    return (int)t.numel();
}
void refnets_param_data(RefNets* r, int i, float* dst) { torch::NoGradGuard g; out(r->flat[i].second, dst); }

// checkpoint files as DecisionNetHandler::save_models / load_models write and read them
// (torch::save / torch::load of one module, decisionNetHandler.cpp:726-757); net = index into `all`
int refnets_save(RefNets* r, int net, const char* path) {
    try {
        auto* m = r->all[net].second;
        torch::serialize::OutputArchive archive;
        m->save(archive);
        archive.save_to(path);
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "refnets_save: %s\n", e.what()); return -1; }
}
int refnets_load(RefNets* r, int net, const char* path) {
    try {
        auto* m = r->all[net].second;
        torch::serialize::InputArchive archive;
        archive.load_from(path);
        m->load(archive);
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "refnets_load: %s\n", e.what()); return -1; }
}
const char* refnets_net_name(RefNets* r, int net) { return r->all[net].first.c_str(); }

// forwards (batch-1, exactly as DecisionNetHandler calls them)
void refnets_encode(RefNets* r, int job, const float* x, int n, float* y) {
    torch::NoGradGuard g;
    auto enc = job ? r->jobOfferEncoder : r->offerEncoder;
    out(enc->forward(t2(x, n, job ? 2 : r->G + 1)), y);
}
// which: 0 purchaseNet, 1 firmPurchaseNet, 2 laborSearchNet
void refnets_purchase(RefNets* r, int which, const float* offerEnc, const float* params, int nparams, float budget,
                      float labor, const float* inv, float* y) {
    torch::NoGradGuard g;
    auto net = which == 0 ? r->purchaseNet : which == 1 ? r->firmPurchaseNet : r->laborSearchNet;
    out(net->forward(t2(offerEnc, r->S, r->enc), t1(params, nparams), torch::tensor({budget}), torch::tensor({labor}),
                     t1(inv, r->G)), y);
}
// which: 0 consumptionNet, 1 productionNet ; y[G][2]
void refnets_consumption(RefNets* r, int which, const float* params, int nparams, float money, float labor,
                         const float* inv, float* y) {
    torch::NoGradGuard g;
    auto net = which == 0 ? r->consumptionNet : r->productionNet;
    out(net->forward(t1(params, nparams), torch::tensor({money}), torch::tensor({labor}), t1(inv, r->G)), y);
}
void refnets_offer(RefNets* r, const float* offerEnc, const float* params, int nparams, float money, float labor,
                   const float* inv, float* y /*[G][4]*/) {
    torch::NoGradGuard g;
    out(r->offerNet->forward(t2(offerEnc, r->S, r->enc), t1(params, nparams), torch::tensor({money}),
                             torch::tensor({labor}), t1(inv, r->G)), y);
}
void refnets_joboffer(RefNets* r, const float* jobEnc, const float* params, int nparams, float money, float labor,
                      const float* inv, float* y /*[4]*/) {
    torch::NoGradGuard g;
    out(r->jobOfferNet->forward(t2(jobEnc, r->S, r->enc), t1(params, nparams), torch::tensor({money}),
                                torch::tensor({labor}), t1(inv, r->G)), y);
}
// which: 0 valueNet, 1 firmValueNet
void refnets_value(RefNets* r, int which, const float* offerEnc, const float* jobEnc, const float* params, int nparams,
                   float money, float labor, const float* inv, float* y /*[1]*/) {
    torch::NoGradGuard g;
    auto net = which == 0 ? r->valueNet : r->firmValueNet;
    out(net->forward(t2(offerEnc, r->S, r->enc), t2(jobEnc, r->S, r->enc), t1(params, nparams), torch::tensor({money}),
                     torch::tensor({labor}), t1(inv, r->G)), y);
}

}  // extern "C"
