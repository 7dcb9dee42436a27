// bank.hpp -- read-set input surface of the Simka hot path, host side.
//
// Same names and calling conventions as the GATB-core classes simka uses to read its samples
// (IBank GC/bank/api/IBank.hpp:78-146, Iterator<T> GC/tools/designpattern/api/Iterator.hpp:117-207,
// Sequence GC/bank/api/Sequence.hpp:70-111, Bank::open GC/bank/impl/Bank.cpp:143-160, BankFasta
// GC/bank/impl/BankFasta.cpp:391-572, BankAlbum GC/bank/impl/BankAlbum.cpp:310-348), re-implemented on
// plain C++17 + zlib: no GATB dependency, no intrusive ref-counting (callers own what they `new`).
// On top of the one-read-at-a-time iterator there is a batch reader (FastxParser::appendReads) that fills
// the concatenated-ASCII + offsets layout simka_pack_reads / simka_gpu_push_ascii take.
#pragma once
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <cctype>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace simka_b200 {
namespace bank {

struct Exception : std::runtime_error {
    explicit Exception(const std::string &m) : std::runtime_error(m) {}
};

// ---- Sequence (GC/bank/api/Sequence.hpp:70-111): data is ASCII, as with Data::ASCII ----------------
class Sequence {
public:
    char *getDataBuffer() { return data.empty() ? nullptr : &data[0]; }
    const char *getDataBuffer() const { return data.data(); }
    size_t getDataSize() const { return data.size(); }
    const std::string &getData() const { return data; }
    const std::string &getComment() const { return comment; }
    const std::string &getQuality() const { return quality; }
    size_t getIndex() const { return index; }
    std::string toString() const { return data; }
    std::string comment, data, quality;
    size_t index = 0;
};

// ---- Iterator<T> (GC/tools/designpattern/api/Iterator.hpp:117-207) ----------------------------------
template <class Item>
class Iterator {
public:
    virtual ~Iterator() {}
    virtual void first() = 0;
    virtual void next() = 0;
    virtual bool isDone() = 0;
    virtual Item &item() = 0;
    Item *operator->() { return &item(); }
    Item &operator*() { return item(); }
    // sub-iterators of a composite bank (album); a plain iterator is its own composition
    virtual std::vector<Iterator<Item> *> getComposition() { return std::vector<Iterator<Item> *>(1, this); }
    virtual void finalize() {}
};

// ---- streaming FASTA / FASTQ (+gz) parser ---------------------------------------------------------------
// kseq-style grammar, as BankFasta: a record starts at a line beginning with '>' or '@'; sequence lines are
// concatenated until a line starting with '>', '@' or '+'; after '+' quality characters are consumed until
// at least as many as bases were read.  Multi-line FASTA, FASTQ, CRLF and gz are handled.
class FastxParser {
public:
    explicit FastxParser(const std::string &path) : path_(path) {
        // plain files are read with read(2) into the parse buffer; only gzip streams (magic 1f 8b) go through zlib
        raw_ = ::open(path.c_str(), O_RDONLY);
        if (raw_ < 0) throw Exception("cannot open " + path);
        unsigned char magic[2] = {0, 0};
        const bool gz = ::pread(raw_, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
        if (gz) {
            ::close(raw_);
            raw_ = -1;
            f_ = gzopen(path.c_str(), "rb");
            if (!f_) throw Exception("cannot open " + path);
            gzbuffer(f_, 1 << 20);
        }
        buf_.resize(1 << 21);
    }
    ~FastxParser() {
        if (f_) gzclose(f_);
        if (raw_ >= 0) ::close(raw_);
    }
    FastxParser(const FastxParser &) = delete;
    FastxParser &operator=(const FastxParser &) = delete;

    void rewind() {
        if (f_) gzrewind(f_);
        else ::lseek(raw_, 0, SEEK_SET);
        beg_ = end_ = 0;
        eof_ = false;
        last_ = 0;
        index_ = 0;
    }

    // Appends the next read's bases to `data` (nothing else is touched).  Optionally returns header / quality.
    // Returns false at end of file.
    bool nextRead(std::string &data, std::string *comment = nullptr, std::string *quality = nullptr) {
        int c;
        if (last_ == 0) {
            while ((c = getc_()) != -1 && c != '>' && c != '@') {}
            if (c == -1) return false;
            last_ = c;
        }
        if (comment) {
            comment->clear();
            appendLine_(*comment);
            if (!comment->empty() && comment->back() == '\r') comment->pop_back();
        } else {
            skipLine_();
        }
        const size_t start = data.size();
        for (;;) {  // sequence lines up to the next line that starts with '>', '+' or '@'
            if (beg_ >= end_ && !fill_()) { c = -1; break; }
            c = (unsigned char)buf_[beg_];
            if (c == '>' || c == '+' || c == '@') { beg_++; break; }
            if (c == '\n') { beg_++; continue; }
            appendLine_(data);
            if (data.size() - start > 1 && data.back() == '\r') data.pop_back();
        }
        last_ = (c == '>' || c == '@') ? c : 0;
        if (quality) quality->clear();
        if (c == '+') {
            const size_t rlen = data.size() - start;
            size_t qlen = 0;
            c = skipLine_() ? '\n' : -1;
            while (c != -1) {
                size_t got = 0;
                c = consumeLine_(quality, got) ? '\n' : -1;
                qlen += got;
                const bool any = got > 0;
                if (c == -1 && !any) break;
                if (qlen >= rlen) break;
            }
            last_ = 0;
        }
        index_++;
        return true;
    }
    size_t nbReadsSoFar() const { return index_; }
    const std::string &path() const { return path_; }

private:
    bool fill_() {
        if (eof_) return false;
        long n;
        if (f_) n = gzread(f_, &buf_[0], (unsigned)buf_.size());
        else do { n = (long)::read(raw_, &buf_[0], buf_.size()); } while (n < 0 && errno == EINTR);
        if (n < 0) throw Exception("read error in " + path_);
        beg_ = 0;
        end_ = (size_t)n;
        if (n == 0) eof_ = true;
        return n > 0;
    }
    int getc_() {
        if (beg_ >= end_ && !fill_()) return -1;
        return (unsigned char)buf_[beg_++];
    }
    // drops the rest of the current line including its newline; false at end of file before a newline
    bool skipLine_() {
        for (;;) {
            if (beg_ >= end_ && !fill_()) return false;
            const char *p = &buf_[beg_];
            const char *nl = (const char *)memchr(p, '\n', end_ - beg_);
            if (nl) { beg_ += (size_t)(nl - p) + 1; return true; }
            beg_ = end_;
        }
    }
    // consumes the rest of the current line: `got` = its characters (newline excluded), appended to `out` if given;
    // false at end of file before a newline
    bool consumeLine_(std::string *out, size_t &got) {
        got = 0;
        for (;;) {
            if (beg_ >= end_ && !fill_()) return false;
            const char *p = &buf_[beg_];
            const char *nl = (const char *)memchr(p, '\n', end_ - beg_);
            const size_t n = nl ? (size_t)(nl - p) : end_ - beg_;
            if (out) out->append(p, n);
            got += n;
            beg_ += n + (nl ? 1 : 0);
            if (nl) return true;
        }
    }
    // rest of the current line (without the newline) appended to `out`
    void appendLine_(std::string &out) {
        for (;;) {
            if (beg_ >= end_ && !fill_()) return;
            const char *p = &buf_[beg_];
            const char *nl = (const char *)memchr(p, '\n', end_ - beg_);
            if (nl) {
                out.append(p, nl - p);
                beg_ += (size_t)(nl - p) + 1;
                return;
            }
            out.append(p, end_ - beg_);
            beg_ = end_;
        }
    }
    std::string path_;
    gzFile f_ = nullptr;
    int raw_ = -1;
    std::string buf_;
    size_t beg_ = 0, end_ = 0;
    bool eof_ = false;
    int last_ = 0;
    size_t index_ = 0;
};

// ---- IBank (GC/bank/api/IBank.hpp:78-146) ------------------------------------------------------------------
class IBank {
public:
    virtual ~IBank() {}
    virtual std::string getId() = 0;
    virtual Iterator<Sequence> *iterator() = 0;  // caller deletes
    virtual int64_t getNbItems() { return -1; }
    // as the reference: from the head of the file, extrapolated (BankFasta.cpp:728-775)
    virtual void estimate(uint64_t &number, uint64_t &totalSize, uint64_t &maxSize) = 0;
    int64_t estimateNbItems() { uint64_t n, t, m; estimate(n, t, m); return (int64_t)n; }
    virtual size_t getCompositionNb() { return 1; }
    virtual void finalize() {}
    // the files behind this bank, in iteration order
    virtual std::vector<std::string> files() = 0;
};

class BankFasta : public IBank {
public:
    explicit BankFasta(const std::string &path) : path_(path) {}
    static const char *name() { return "fasta"; }
    std::string getId() override { return path_; }
    std::vector<std::string> files() override { return std::vector<std::string>(1, path_); }
    // BankFasta::Iterator::estimate (GC/bank/impl/BankFasta.cpp:728-775): the sequences at the head of the file are read (until
    // more than the threshold of 50000 have been seen) and the counts are extrapolated linearly from the position of the
    // reader -- what gzread has delivered so far, in buffers of 256 KB -- to the size of the file, a gz file being taken for
    // four times its compressed size (:167-188).  Exact for a plain file that ends before the threshold; NOT the number of
    // reads of a gz file.  `simka -max-reads 0` derives its limit from these numbers (S/core/SimkaAlgorithm.cpp:397-430).
    void estimate(uint64_t &number, uint64_t &totalSize, uint64_t &maxSize) override {
        HeadReader r(path_);
        number = totalSize = maxSize = 0;
        uint64_t len = 0;
        while (r.nextSequence(len) && number <= estimateThreshold_) {
            number++;
            if (len > maxSize) maxSize = len;
            totalSize += len;
        }
        const uint64_t actual = r.delivered();
        if (actual > 0) {
            const uint64_t size = estimatedFileSize(path_);
            number = (number * size) / actual;
            totalSize = (uint64_t)((float)totalSize * ((float)size / (float)actual));
        }
    }
    void setEstimateThreshold(uint64_t n) { estimateThreshold_ = n; }
    // file size as the reference sees it: a name whose FIRST "gz" is its last two characters counts four times its size
    static uint64_t estimatedFileSize(const std::string &path) {
        gzFile f = gzopen(path.c_str(), "rb");  // (existence check with the same error as the readers)
        if (!f) throw Exception("Unable to open bank '" + path + "'");
        gzclose(f);
        std::ifstream in(path.c_str(), std::ios::binary | std::ios::ate);
        const uint64_t sz = in ? (uint64_t)in.tellg() : 0;
        const char *name = path.c_str();
        const char *gz = strstr(name, "gz");
        return (gz != nullptr && gz == name + strlen(name) - 2) ? sz * 4 : sz;
    }
    // the reference's buffered kseq-style reader (BankFasta.cpp:391-572), reduced to what estimate() needs: where records end
    // and how many 256 KB buffers have been fetched
    class HeadReader {
    public:
        explicit HeadReader(const std::string &path) : buf_(BUFFER_SIZE) {
            f_ = gzopen(path.c_str(), "rb");
            if (!f_) throw Exception("Unable to open bank '" + path + "'");
        }
        ~HeadReader() { if (f_) gzclose(f_); }
        HeadReader(const HeadReader &) = delete;
        HeadReader &operator=(const HeadReader &) = delete;
        uint64_t delivered() const { return delivered_; }  // gztell of the reference's stream
        bool nextSequence(uint64_t &length) {  // get_next_seq_from_file (:485-572); length = number of sequence characters
            int c;
            if (last_char_ == 0) {
                while ((c = getc_()) != -1 && c != '>' && c != '@') {}
                if (c == -1) return false;
                last_char_ = c;
            }
            uint64_t hlen = 0;
            char dret = 0;
            if (gets_(hlen, &dret, false) < 0) return false;
            if (dret != '\n') { uint64_t rest = 0; gets_(rest, nullptr, true); }
            uint64_t rlen = 0;
            while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
                if (c == '\n') continue;
                rlen++;
                last_of_string_ = (char)c;
                gets_(rlen, nullptr, true);
            }
            if (c == '>' || c == '@') last_char_ = c;
            if (c == '+') {
                while ((c = getc_()) != -1 && c != '\n') {}
                uint64_t qlen = 0;
                last_of_string_ = 0;
                while (gets_(qlen, nullptr, true) >= 0 && qlen < rlen) {}
                last_char_ = 0;
            }
            length = rlen;
            return true;
        }
    private:
        static const int BUFFER_SIZE = 256 * 1024;
        bool rebuffer_() {
            if (eof_) return false;
            start_ = 0;
            const int n = gzread(f_, buf_.data(), BUFFER_SIZE);
            end_ = n > 0 ? (size_t)n : 0;
            delivered_ += end_;
            if (end_ < (size_t)BUFFER_SIZE) eof_ = true;
            return end_ != 0;
        }
        int getc_() {
            if (start_ >= end_) if (!rebuffer_()) return -1;
            return (int)(signed char)buf_[start_++];
        }
        // buffered_gets (:421-475) in append mode on a string of which only the length (and last character) is kept
        int64_t gets_(uint64_t &length, char *dret, bool allow_spaces) {
            if (dret) *dret = 0;
            if (start_ >= end_ && eof_) return -1;
            for (;;) {
                if (start_ >= end_) if (!rebuffer_()) break;
                size_t i = start_;
                if (allow_spaces) { while (i < end_ && buf_[i] != '\n') i++; }
                else { while (i < end_ && !isspace((unsigned char)buf_[i])) i++; }
                if (i > start_) last_of_string_ = (char)buf_[i - 1];
                length += i - start_;
                start_ = i + 1;
                if (i < end_) { if (dret) *dret = (char)buf_[i]; break; }
            }
            if (allow_spaces && length > 1 && last_of_string_ == '\r') { length--; last_of_string_ = 0; }
            return (int64_t)length;
        }
        gzFile f_ = nullptr;
        std::vector<unsigned char> buf_;
        size_t start_ = 0, end_ = 0;
        bool eof_ = false;
        uint64_t delivered_ = 0;
        int last_char_ = 0;
        char last_of_string_ = 0;
    };
    class It : public Iterator<Sequence> {
    public:
        explicit It(const std::string &path) : parser_(path) {}
        void first() override { parser_.rewind(); done_ = false; next(); }
        void next() override {
            seq_.data.clear();
            seq_.index = parser_.nbReadsSoFar();
            done_ = !parser_.nextRead(seq_.data, &seq_.comment, &seq_.quality);
        }
        bool isDone() override { return done_; }
        Sequence &item() override { return seq_; }
    private:
        FastxParser parser_;
        Sequence seq_;
        bool done_ = true;
    };
    Iterator<Sequence> *iterator() override { return new It(path_); }
private:
    std::string path_;
    uint64_t estimateThreshold_ = 50000;  // AbstractBank.hpp:54
};

// ---- BankAlbum: a text file with one bank uri per line (relative to the album's directory) --------------------
class BankAlbum : public IBank {
public:
    explicit BankAlbum(const std::string &path) : path_(path) {
        std::ifstream in(path.c_str());
        if (!in) throw Exception("cannot open album " + path);
        std::string dir = path.find('/') == std::string::npos ? std::string(".") : path.substr(0, path.rfind('/'));
        std::string line;
        while (std::getline(in, line)) {
            while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
            size_t b = 0;
            while (b < line.size() && (line[b] == ' ' || line[b] == '\t')) b++;
            line = line.substr(b);
            if (line.empty() || line[0] == '#') continue;
            files_.push_back(line[0] == '/' ? line : dir + "/" + line);
        }
    }
    static const char *name() { return "album"; }
    std::string getId() override { return path_; }
    std::vector<std::string> files() override { return files_; }
    size_t getCompositionNb() override { return files_.size(); }
    void estimate(uint64_t &number, uint64_t &totalSize, uint64_t &maxSize) override {
        number = totalSize = maxSize = 0;
        for (const std::string &f : files_) {
            uint64_t n, t, m;
            BankFasta(f).estimate(n, t, m);
            number += n; totalSize += t; if (m > maxSize) maxSize = m;
        }
    }
    // composite iterator: the files one after the other; getComposition() exposes the per-file iterators
    class It : public Iterator<Sequence> {
    public:
        explicit It(const std::vector<std::string> &files) { for (const std::string &f : files) its_.emplace_back(new BankFasta::It(f)); }
        void first() override { cur_ = 0; if (!its_.empty()) its_[0]->first(); skip_(); }
        void next() override { its_[cur_]->next(); skip_(); }
        bool isDone() override { return cur_ >= its_.size(); }
        Sequence &item() override { return its_[cur_]->item(); }
        std::vector<Iterator<Sequence> *> getComposition() override {
            std::vector<Iterator<Sequence> *> v;
            for (auto &p : its_) v.push_back(p.get());
            return v;
        }
    private:
        void skip_() { while (cur_ < its_.size() && its_[cur_]->isDone()) { cur_++; if (cur_ < its_.size()) its_[cur_]->first(); } }
        std::vector<std::unique_ptr<BankFasta::It>> its_;
        size_t cur_ = 0;
    };
    Iterator<Sequence> *iterator() override { return new It(files_); }
private:
    std::string path_;
    std::vector<std::string> files_;
};

// ---- Bank::open (GC/bank/impl/Bank.cpp:143-160): fasta/fastq if the content starts a record, else album ----------
struct Bank {
    static IBank *open(const std::string &uri) {
        gzFile f = gzopen(uri.c_str(), "rb");
        if (!f) throw Exception("Unable to open bank '" + uri + "'");
        int c;
        do { c = gzgetc(f); } while (c == ' ' || c == '\n' || c == '\r' || c == '\t');
        gzclose(f);
        if (c == '>' || c == '@') return new BankFasta(uri);
        return new BankAlbum(uri);
    }
};

}  // namespace bank
}  // namespace simka_b200
