// kmer_model.hpp -- host-side k-mer model with the surface of GATB-core's Kmer<span>::Model*.
//
// Mirrors the API simka touches (src/minikc/MiniKC.hpp:204-252, src/simkaMin/SimkaMinCount.hpp:219-247):
//   Kmer<>::ModelCanonical model(k);  Kmer<>::ModelCanonical::Iterator it(model);
//   it.setData(seq.getData());  for (it.first(); !it.isDone(); it.next()) it->value().getVal();
//   model.build(data, kmers);  kmers[i].isValid();  kmers[i].value();  kmers[i].forward();  kmers[i].revcomp();
//   Kmer<>::ModelMinimizer<ModelCanonical> mm(k, m);  mm.getMinimizerValue(kmer);
// with the arithmetic of GC/kmer/impl/Model.hpp:637-657,726-765,858-884 (A=0 C=1 T=2 G=3 = (c>>1)&3, first base
// most significant, canonical = min(forward, reverse complement) as unsigned integers, a k-mer is invalid when
// any of its k characters is outside ACGTacgt) and :1010-1070,1220-1287 (minimizer = smallest canonical m-mer that
// does not contain "AA" after its first base; default 4^m-1).  Only span 32 (k <= 31, one u64): the device
// kernels share the same encoding through kmer_device.cuh.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../kmer_device.cuh"

namespace simka_b200 {

inline bool nt_valid(unsigned char c) {
    c &= 0xDF;
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}
inline unsigned nt_code(unsigned char c) { return (c >> 1) & 3u; }

template <size_t span = 32>
struct Kmer {
    static_assert(span == 32, "only span 32 (k <= 31) is implemented");

    // the integer type of a k-mer (LargeInt<1>, GC/tools/math/LargeInt1.pri:42-85)
    class Type {
    public:
        Type(uint64_t v = 0) : v_(v) {}
        uint64_t getVal() const { return v_; }
        void setVal(uint64_t v) { v_ = v; }
        bool operator<(const Type &o) const { return v_ < o.v_; }
        bool operator<=(const Type &o) const { return v_ <= o.v_; }
        bool operator==(const Type &o) const { return v_ == o.v_; }
        bool operator!=(const Type &o) const { return v_ != o.v_; }
        Type operator<<(int s) const { return Type(v_ << s); }
        Type operator>>(int s) const { return Type(v_ >> s); }
        Type operator&(const Type &o) const { return Type(v_ & o.v_); }
        Type operator|(const Type &o) const { return Type(v_ | o.v_); }
        unsigned operator[](size_t i) const { return (unsigned)((v_ >> (2 * i)) & 3); }  // i-th base from the END
        std::string toString(size_t k) const {
            static const char bin2nt[4] = {'A', 'C', 'T', 'G'};
            std::string s(k, 'A');
            for (size_t i = 0; i < k; i++) s[k - 1 - i] = bin2nt[(v_ >> (2 * i)) & 3];
            return s;
        }
    private:
        uint64_t v_;
    };

    class KmerDirect {
    public:
        const Type &value() const { return value_; }
        const Type &forward() const { return value_; }
        bool isValid() const { return valid_; }
        Type value_;
        bool valid_ = true;
    };

    class KmerCanonical {
    public:
        const Type &value() const { return choice_ ? rev_ : fwd_; }
        const Type &forward() const { return fwd_; }
        const Type &revcomp() const { return rev_; }
        bool which() const { return !choice_; }  // true if the forward strand is the canonical one
        bool isValid() const { return valid_; }
        Type fwd_, rev_;
        bool choice_ = false, valid_ = true;
    };

    template <bool CANONICAL>
    class ModelBase {
    public:
        typedef typename std::conditional<CANONICAL, KmerCanonical, KmerDirect>::type Kmer_t;
        explicit ModelBase(size_t k = span - 1) : k_(k) {
            if (k < 1 || k >= span) throw std::runtime_error("Type '" + std::to_string(span) + "' doesn't support kmer size " + std::to_string(k));
            mask_ = k_ == 32 ? ~0ull : ((1ull << (2 * k_)) - 1);
        }
        size_t getKmerSize() const { return k_; }
        uint64_t getKmerMax() const { return mask_; }
        std::string toString(const Type &t) const { return t.toString(k_); }
        Type reverse(const Type &t) const { return Type(simka::revcomp(t.getVal(), (int)k_)); }

        // all k-mers of a sequence, including the invalid ones (flagged) -- Model::build (Model.hpp:499-513)
        bool build(const std::string &data, std::vector<Kmer_t> &out) const { return build(data.data(), data.size(), out); }
        bool build(const char *seq, size_t len, std::vector<Kmer_t> &out) const {
            out.clear();
            if (len < k_) return false;
            out.resize(len - k_ + 1);
            uint64_t fwd = 0, rev = 0;
            long bad = 0;  // > 0 while a bad character is inside the window
            for (size_t i = 0; i < len; i++) {
                const unsigned c = nt_code((unsigned char)seq[i]);
                fwd = ((fwd << 2) | c) & mask_;
                rev = (rev >> 2) | ((uint64_t)(c ^ 2u) << (2 * (k_ - 1)));
                if (!nt_valid((unsigned char)seq[i])) bad = (long)k_; else if (bad > 0) bad--;
                if (i + 1 >= k_) set_(out[i + 1 - k_], fwd, rev, bad == 0);
            }
            return true;
        }

        // Model::Iterator (Model.hpp:520-616): iterates the k-mers of the sequence given with setData
        class Iterator {
        public:
            explicit Iterator(const ModelBase &m) : model_(m) {}
            void setData(const std::string &d) { model_.build(d, kmers_); pos_ = 0; }
            void setData(const char *p, size_t n) { model_.build(p, n, kmers_); pos_ = 0; }
            void first() { pos_ = 0; }
            void next() { pos_++; }
            bool isDone() const { return pos_ >= kmers_.size(); }
            Kmer_t &item() { return kmers_[pos_]; }
            Kmer_t *operator->() { return &kmers_[pos_]; }
            Kmer_t &operator*() { return kmers_[pos_]; }
        private:
            const ModelBase &model_;
            std::vector<Kmer_t> kmers_;
            size_t pos_ = 0;
        };

    private:
        static void set_(KmerDirect &o, uint64_t fwd, uint64_t, bool valid) { o.value_ = Type(fwd); o.valid_ = valid; }
        static void set_(KmerCanonical &o, uint64_t fwd, uint64_t rev, bool valid) {
            o.fwd_ = Type(fwd); o.rev_ = Type(rev); o.choice_ = rev < fwd; o.valid_ = valid;
        }
        size_t k_;
        uint64_t mask_;
    };

    typedef ModelBase<false> ModelDirect;
    typedef ModelBase<true> ModelCanonical;

    // ModelMinimizer<ModelCanonical> (Model.hpp:1010-1148): only what simka calls, getMinimizerValue
    template <class Model = ModelCanonical>
    class ModelMinimizer : public Model {
    public:
        ModelMinimizer(size_t k, size_t m) : Model(k), m_(m) {
            if (m < 2 || m > 12 || m > k) throw std::runtime_error("bad minimizer size");
            const uint32_t n = 1u << (2 * m);
            lut_.resize(n);
            for (uint32_t v = 0; v < n; v++) {
                uint32_t rc = (uint32_t)simka::revcomp(v, (int)m);
                uint32_t mm = rc < v ? rc : v;
                lut_[v] = allowed_(mm) ? mm : n - 1;
            }
        }
        size_t getMmersSize() const { return m_; }
        uint64_t getMinimizerValue(const Type &kmer) const {
            const uint32_t mmask = (1u << (2 * m_)) - 1;
            uint32_t best = mmask;
            uint64_t val = kmer.getVal();
            for (long idx = (long)this->getKmerSize() - (long)m_; idx >= 0; idx--) {
                const uint32_t mm = lut_[(uint32_t)(val & mmask)];
                if (mm < best) best = mm;
                val >>= 2;
            }
            return best;
        }
    private:
        bool allowed_(uint32_t mmer) const {  // no "AA" except at the first position (Model.hpp:1220-1251)
            const uint64_t mmask_m1 = (1ull << ((m_ - 2) * 2)) - 1;
            const uint64_t mask_ma1 = 0x5555555555555555ull & mmask_m1;
            uint64_t a1 = mmer;
            a1 = ~(a1 | (a1 >> 2));
            a1 = ((a1 >> 1) & a1) & mask_ma1;
            return a1 == 0;
        }
        size_t m_;
        std::vector<uint32_t> lut_;
    };
};

}  // namespace simka_b200
