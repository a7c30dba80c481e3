// Protocol-level C ABI: simpleworks::marlin's surface (reference src/marlin/mod.rs:33-94) on the
// CUDA engine.  Host orchestration (transcript, AHP rounds, KZG bookkeeping) is marlin/marlin.hpp;
// every NTT, MSM and fixed-base table goes to the kernels of this library -- there is no CPU
// engine in libswb200.
#include "ctx.hpp"
#include "marlin/c_api_impl.hpp"

using namespace swb;
using namespace swb::marlin;

namespace {

struct GpuEngine {
    swb_ctx* c;
    void fail(const char* what) { throw MarlinError(std::string(what) + ": " + swb_last_error(c)); }
    void ntt(Fr* v, uint32_t log_n, bool inverse, bool coset) {
        if (swb_ntt_fr(c, reinterpret_cast<swb_fr*>(v), log_n, inverse, coset) != SWB_OK) fail("ntt");
    }
    void* bases_from_powers(const G1Point& g, const Fr& beta, size_t n) {
        swb_g1_jacobian gj;
        memset(&gj, 0, sizeof gj);
        Fq one = Fq::one();
        if (g.infinity) memcpy(gj.y.l, one.l, 48);
        else {
            memcpy(gj.x.l, g.x.l, 48);
            memcpy(gj.y.l, g.y.l, 48);
            memcpy(gj.z.l, one.l, 48);
        }
        swb_fr b;
        memcpy(b.l, beta.l, 32);
        swb_bases* out = nullptr;
        if (swb_bases_from_powers(c, &gj, &b, n, &out) != SWB_OK) fail("bases_from_powers");
        return out;
    }
    void export_bases(void* h, size_t offset, size_t n, G1Point* out) {
        std::vector<swb_g1_affine> tmp(n);
        if (swb_bases_export(c, static_cast<swb_bases*>(h), offset, n, tmp.data()) != SWB_OK) fail("bases_export");
        for (size_t i = 0; i < n; i++) {
            out[i].infinity = tmp[i].infinity != 0;
            memcpy(out[i].x.l, tmp[i].x.l, 48);
            memcpy(out[i].y.l, tmp[i].y.l, 48);
        }
    }
    void free_bases(void* h) { swb_bases_free(static_cast<swb_bases*>(h)); }
    G1Point msm(void* h, size_t offset, const Fr* scalars_mont, size_t n) {
        swb_g1_jacobian out;
        if (swb_msm_g1_fr(c, static_cast<swb_bases*>(h), offset, reinterpret_cast<const swb_fr*>(scalars_mont), n, &out) != SWB_OK)
            fail("msm");
        G1Point p = G1Point::identity();
        Fq z;
        memcpy(z.l, out.z.l, 48);
        if (!z.is_zero()) {                 // the library returns Z = 1
            p.infinity = false;
            memcpy(p.x.l, out.x.l, 48);
            memcpy(p.y.l, out.y.l, 48);
        }
        return p;
    }
};
using Api = MarlinApi<GpuEngine>;

}  // namespace

struct swb_rng { RngHandle h; };
struct swb_r1cs { R1csHandle* h; };
struct swb_srs { GpuEngine eng; SrsHandle<GpuEngine>* h; };
struct swb_pk { PkHandle<GpuEngine>* h; };
struct swb_vk { VkHandle<GpuEngine>* h; };

extern "C" {

swb_rng* swb_rng_test_rng(void) {
    auto* r = new swb_rng();
    r->h.rng = test_rng();
    return r;
}
uint64_t swb_rng_next_u64(swb_rng* r) { return r->h.rng.next_u64(); }
void swb_rng_free(swb_rng* r) { delete r; }

swb_r1cs* swb_r1cs_new(size_t num_instance, size_t num_witness) { return new swb_r1cs{r1cs_new(num_instance, num_witness)}; }
swb_r1cs* swb_r1cs_builtin(int kind, size_t size, uint64_t v0, uint64_t v1) {
    R1csHandle* h = r1cs_builtin(kind, size, v0, v1);
    return h ? new swb_r1cs{h} : nullptr;
}
int swb_r1cs_add_constraint(swb_r1cs* cs, const swb_fr* a_coef, const uint32_t* a_col, size_t na, const swb_fr* b_coef,
                            const uint32_t* b_col, size_t nb, const swb_fr* c_coef, const uint32_t* c_col, size_t nc) {
    if (!cs) return SWB_EARG;
    return r1cs_add_constraint(cs->h, (const uint64_t*)a_coef, a_col, na, (const uint64_t*)b_coef, b_col, nb, (const uint64_t*)c_coef,
                               c_col, nc) ? SWB_EARG : SWB_OK;
}
int swb_r1cs_set_assignment(swb_r1cs* cs, const swb_fr* instance, size_t ni, const swb_fr* witness, size_t nw) {
    if (!cs) return SWB_EARG;
    return r1cs_set_assignment(cs->h, (const uint64_t*)instance, ni, (const uint64_t*)witness, nw) ? SWB_EARG : SWB_OK;
}
int swb_r1cs_is_satisfied(const swb_r1cs* cs) { return cs && cs->h->cs.is_satisfied() ? 1 : 0; }
void swb_r1cs_free(swb_r1cs* cs) {
    if (!cs) return;
    delete cs->h;
    delete cs;
}

int swb_marlin_universal_setup(swb_ctx* c, size_t nc, size_t nv, size_t nnz, swb_rng* rng, swb_srs** out) {
    if (!c || !rng || !out) return SWB_EARG;
    auto* s = new swb_srs{GpuEngine{c}, nullptr};
    std::string err;
    int rc = Api::setup(s->eng, nc, nv, nnz, &rng->h, &s->h, &err);
    if (rc) {
        delete s;
        return swb::set_err(c, SWB_EINTERNAL, "universal_setup: %s", err.c_str());
    }
    *out = s;
    return SWB_OK;
}
size_t swb_srs_max_degree(const swb_srs* s) { return s ? s->h->srs->max_degree : 0; }
void swb_srs_free(swb_srs* s) {
    if (!s) return;
    delete s->h;
    delete s;
}
int swb_marlin_index(swb_ctx* c, const swb_srs* srs, const swb_r1cs* cs, swb_pk** pk, swb_vk** vk) {
    if (!c || !srs || !cs || !pk || !vk) return SWB_EARG;
    std::string err;
    auto* p = new swb_pk{nullptr};
    auto* v = new swb_vk{nullptr};
    GpuEngine eng{c};
    int rc = Api::index(eng, srs->h, cs->h, &p->h, &v->h, &err);
    if (rc) {
        delete p;
        delete v;
        return swb::set_err(c, SWB_EINTERNAL, "index: %s", err.c_str());
    }
    *pk = p;
    *vk = v;
    return SWB_OK;
}
void swb_pk_free(swb_pk* p) {
    if (!p) return;
    delete p->h;
    delete p;
}
void swb_vk_free(swb_vk* v) {
    if (!v) return;
    delete v->h;
    delete v;
}
int swb_marlin_prove(swb_ctx* c, const swb_pk* pk, const swb_r1cs* cs, swb_rng* rng, uint8_t** proof, size_t* len) {
    if (!c || !pk || !cs || !rng || !proof || !len) return SWB_EARG;
    std::string err;
    GpuEngine eng{c};
    int rc = Api::prove(eng, pk->h, cs->h, &rng->h, proof, len, &err);
    if (rc) return swb::set_err(c, SWB_EINTERNAL, "prove: %s", err.c_str());
    return SWB_OK;
}
int swb_marlin_verify(swb_ctx* c, const swb_vk* vk, const swb_fr* public_inputs, size_t n, const uint8_t* proof, size_t len, int* ok) {
    if (!vk || !ok || (!public_inputs && n) || !proof) return SWB_EARG;
    std::string err;
    int rc = Api::verify(vk->h, (const uint64_t*)public_inputs, n, proof, len, ok, &err);
    if (rc) return swb::set_err(c, SWB_EINTERNAL, "verify: %s", err.c_str());
    return SWB_OK;
}
void swb_bytes_free(uint8_t* p) { free(p); }

}  // extern "C"
