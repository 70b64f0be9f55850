#include "seqio.hpp"

#include <cstring>
#include <stdexcept>

namespace txr
{
SeqReader::SeqReader(const std::string &path)
{
    gz_ = gzopen(path.c_str(), "rb"); // transparently reads plain files as well
    if (gz_)
        gzbuffer(gz_, 1 << 20);
    buf_.resize(1 << 22);
    buf_.resize(0);
}

SeqReader::~SeqReader()
{
    if (gz_)
        gzclose(gz_);
}

int SeqReader::peek()
{
    if (pos_ >= buf_.size())
    {
        if (eof_)
            return -1;
        buf_.resize(1 << 22);
        const int n = gzread(gz_, &buf_[0], (unsigned)buf_.size());
        if (n < 0)
            throw std::runtime_error("read error (corrupt gzip stream?)");
        buf_.resize((size_t)n);
        pos_ = 0;
        if (n == 0)
        {
            eof_ = true;
            return -1;
        }
    }
    return (unsigned char)buf_[pos_];
}

bool SeqReader::getline(std::string &line)
{
    line.clear();
    if (peek() < 0)
        return false;
    while (true)
    {
        const char *start = buf_.data() + pos_;
        const char *nl = static_cast<const char *>(memchr(start, '\n', buf_.size() - pos_));
        if (nl)
        {
            line.append(start, (size_t)(nl - start));
            pos_ += (size_t)(nl - start) + 1;
            break;
        }
        line.append(start, buf_.size() - pos_);
        pos_ = buf_.size();
        if (peek() < 0)
            break;
    }
    if (!line.empty() && line.back() == '\r')
        line.pop_back();
    return true;
}

bool SeqReader::next(std::string &id, std::string &seq)
{
    id.clear();
    seq.clear();
    int c;
    while ((c = peek()) == '\n' || c == '\r') // blank lines between records
        ++pos_;
    if (c < 0)
        return false;
    if (c != '>' && c != '@')
        throw std::runtime_error("sequence file: record does not start with '>' or '@'");
    getline(line_);
    id.assign(line_, 1, std::string::npos);
    if (c == '>')
    {
        while ((c = peek()) >= 0 && c != '>')
        {
            getline(line_);
            seq += line_;
        }
    }
    else
    {
        while ((c = peek()) >= 0 && c != '+')
        {
            getline(line_);
            seq += line_;
        }
        if (c < 0)
            throw std::runtime_error("FASTQ record without a '+' line");
        getline(line_); // '+' line
        size_t q = 0;
        while (q < seq.size() && getline(line_))
            q += line_.size();
        if (q != seq.size())
            throw std::runtime_error("FASTQ quality string length differs from the sequence length");
    }
    return true;
}
} // namespace txr
