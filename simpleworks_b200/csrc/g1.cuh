// BLS12-377 G1 (y^2 = x^3 + 1 over Fq) for the MSM kernels.
//
// Replaces ark_ec::short_weierstrass_jacobian::{GroupAffine, GroupProjective} arithmetic as used
// by VariableBaseMSM (ark-ec 0.3 msm/variable_base.rs; reached from reference
// src/marlin/mod.rs:75,92).  Only affine results are observable upstream (commitments are
// normalised before they are hashed or serialised), so the accumulator representation is free:
// buckets are kept in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), whose mixed addition
// costs 8M + 2S and needs no field inversion.
#pragma once
#include "fp.cuh"

namespace swb {

// device layout of a base point: 96 bytes, x | y, Montgomery; the identity is stored as (0, 0)
// (not on the curve, so it cannot collide with a real point).
struct G1Aff {
    Fq x, y;
    SWB_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct G1Xyzz {
    Fq x, y, zz, zzz;

    static SWB_HD G1Xyzz identity() {
        G1Xyzz r;
        r.x = Fq::zero();
        r.y = Fq::zero();
        r.zz = Fq::zero();
        r.zzz = Fq::zero();
        return r;
    }
    SWB_HD bool is_identity() const { return zz.is_zero(); }

    // 2 * (x2, y2), affine input (mdbl-2008-s-1 with a = 0)
    static SWB_HD G1Xyzz dbl_affine(const Fq& x2, const Fq& y2) {
        G1Xyzz r;
        Fq u = y2.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = x2 * v;
        Fq xx = x2.sqr();
        Fq m = xx.dbl() + xx;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y2;
        r.zz = v;
        r.zzz = w;
        return r;
    }
    // dbl-2008-s-1, a = 0
    SWB_HD G1Xyzz dbl() const {
        if (is_identity()) return *this;
        G1Xyzz r;
        Fq u = y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = x * v;
        Fq xx = x.sqr();
        Fq m = xx.dbl() + xx;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }
    // this += (x2, +-y2)   (madd-2008-s); (x2, y2) must not be the identity
    SWB_HD void add_affine(const Fq& x2, const Fq& y2) {
        if (is_identity()) {
            x = x2;
            y = y2;
            zz = Fq::one();
            zzz = Fq::one();
            return;
        }
        Fq u2 = x2 * zz;
        Fq s2 = y2 * zzz;
        Fq p = u2 - x;
        Fq r = s2 - y;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl_affine(x2, y2);
            else *this = identity();
            return;
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq q = x * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = r * (q - x3) - y * ppp;
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }
    // this += o  (add-2008-s)
    SWB_HD void add(const G1Xyzz& o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        Fq u1 = x * o.zz;
        Fq u2 = o.x * zz;
        Fq s1 = y * o.zzz;
        Fq s2 = o.y * zzz;
        Fq p = u2 - u1;
        Fq r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq q = u1 * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = r * (q - x3) - s1 * ppp;
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }
};

}  // namespace swb
