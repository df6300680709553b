/*
json_min.hpp -- minimal JSON DOM for flame files (host side, parse-once, cold).

Stands in for the reference's wrapper over nlohmann::json (utils/json.hpp:44-105,
utils/json.cpp) which is un-vendored there. Behaviour kept for the flame format:
  - // and block comments are skipped like parse(in,nullptr,true,true)  (json.cpp:16)
  - a number without fraction/exponent is an integer (isInt() true), everything
    else goes through strtod, so each coefficient is the identical double
  - isFloat() is true for any number, like is_number()                  (json.cpp:48-51)
  - a duplicate key keeps the last value
Errors are thrown as ffr::JsonError with the reference's message texts.
*/

#pragma once

#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ffr
{

struct JsonError : public std::runtime_error
{
    explicit JsonError(const std::string& s): std::runtime_error(s) {}
};

class Json
{
public:
    enum Type { Null, Bool, Int, UInt, Float, String, Array, Object };

private:
    Type type = Null;
    bool b = false;
    int64_t i = 0;
    uint64_t u = 0;
    double f = 0.0;
    std::string s;
    std::vector<Json> arr;
    std::map<std::string,Json> obj;

    struct Parser
    {
        const char *p, *end;

        [[noreturn]] void fail(const std::string& msg) const
        {
            throw JsonError("json parse error: " + msg);
        }

        void skipWs()
        {
            for (;;)
            {
                while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r'))
                    ++p;
                if (p+1 < end && p[0] == '/' && p[1] == '/')
                {
                    while (p < end && *p != '\n')
                        ++p;
                }
                else if (p+1 < end && p[0] == '/' && p[1] == '*')
                {
                    p += 2;
                    while (p+1 < end && !(p[0] == '*' && p[1] == '/'))
                        ++p;
                    if (p+1 >= end)
                        fail("unterminated comment");
                    p += 2;
                }
                else
                    return;
            }
        }

        static void appendUtf8(std::string& out, uint32_t cp)
        {
            if (cp < 0x80)
                out += (char)cp;
            else if (cp < 0x800)
            {
                out += (char)(0xC0 | (cp >> 6));
                out += (char)(0x80 | (cp & 0x3F));
            }
            else if (cp < 0x10000)
            {
                out += (char)(0xE0 | (cp >> 12));
                out += (char)(0x80 | ((cp >> 6) & 0x3F));
                out += (char)(0x80 | (cp & 0x3F));
            }
            else
            {
                out += (char)(0xF0 | (cp >> 18));
                out += (char)(0x80 | ((cp >> 12) & 0x3F));
                out += (char)(0x80 | ((cp >> 6) & 0x3F));
                out += (char)(0x80 | (cp & 0x3F));
            }
        }

        uint32_t hex4()
        {
            if (end - p < 4)
                fail("bad \\u escape");
            uint32_t v = 0;
            for (int k = 0; k < 4; ++k)
            {
                char c = *p++;
                v <<= 4;
                if (c >= '0' && c <= '9') v |= c - '0';
                else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
                else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
                else fail("bad \\u escape");
            }
            return v;
        }

        std::string parseString()
        {
            std::string out;
            ++p; // opening quote
            for (;;)
            {
                if (p >= end)
                    fail("unterminated string");
                char c = *p++;
                if (c == '"')
                    return out;
                if (c != '\\')
                {
                    out += c;
                    continue;
                }
                if (p >= end)
                    fail("unterminated string");
                c = *p++;
                switch (c)
                {
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                case '/': out += '/'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;
                case 't': out += '\t'; break;
                case 'u':
                {
                    uint32_t cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6
                        && p[0] == '\\' && p[1] == 'u')
                    {
                        p += 2;
                        uint32_t lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    appendUtf8(out,cp);
                    break;
                }
                default: fail("bad escape");
                }
            }
        }

        Json parseNumber()
        {
            const char *start = p;
            bool is_int = true;
            if (p < end && *p == '-')
                ++p;
            if (p >= end || *p < '0' || *p > '9')
                fail("invalid number");
            while (p < end && *p >= '0' && *p <= '9')
                ++p;
            if (p < end && *p == '.')
            {
                is_int = false;
                ++p;
                if (p >= end || *p < '0' || *p > '9')
                    fail("invalid number");
                while (p < end && *p >= '0' && *p <= '9')
                    ++p;
            }
            if (p < end && (*p == 'e' || *p == 'E'))
            {
                is_int = false;
                ++p;
                if (p < end && (*p == '+' || *p == '-'))
                    ++p;
                if (p >= end || *p < '0' || *p > '9')
                    fail("invalid number");
                while (p < end && *p >= '0' && *p <= '9')
                    ++p;
            }
            std::string tok(start,p);
            Json j;
            if (is_int)
            {
                // same ladder as nlohmann's lexer: signed, then unsigned, then double
                errno = 0;
                char *e = nullptr;
                if (tok[0] == '-')
                {
                    long long v = strtoll(tok.c_str(),&e,10);
                    if (errno == 0)
                    {
                        j.type = Int;
                        j.i = v;
                        return j;
                    }
                }
                else
                {
                    unsigned long long v = strtoull(tok.c_str(),&e,10);
                    if (errno == 0)
                    {
                        if (v <= (unsigned long long)INT64_MAX)
                        {
                            j.type = Int;
                            j.i = (int64_t)v;
                        }
                        else
                        {
                            j.type = UInt;
                            j.u = v;
                        }
                        return j;
                    }
                }
            }
            j.type = Float;
            j.f = strtod(tok.c_str(),nullptr);
            return j;
        }

        Json parseValue(int depth)
        {
            if (depth > 256)
                fail("nesting too deep");
            skipWs();
            if (p >= end)
                fail("unexpected end of input");
            Json j;
            char c = *p;
            if (c == '{')
            {
                ++p;
                j.type = Object;
                skipWs();
                if (p < end && *p == '}')
                {
                    ++p;
                    return j;
                }
                for (;;)
                {
                    skipWs();
                    if (p >= end || *p != '"')
                        fail("expected object key");
                    std::string key = parseString();
                    skipWs();
                    if (p >= end || *p != ':')
                        fail("expected ':'");
                    ++p;
                    j.obj[key] = parseValue(depth+1);
                    skipWs();
                    if (p < end && *p == ',')
                    {
                        ++p;
                        continue;
                    }
                    if (p < end && *p == '}')
                    {
                        ++p;
                        return j;
                    }
                    fail("expected ',' or '}'");
                }
            }
            if (c == '[')
            {
                ++p;
                j.type = Array;
                skipWs();
                if (p < end && *p == ']')
                {
                    ++p;
                    return j;
                }
                for (;;)
                {
                    j.arr.push_back(parseValue(depth+1));
                    skipWs();
                    if (p < end && *p == ',')
                    {
                        ++p;
                        continue;
                    }
                    if (p < end && *p == ']')
                    {
                        ++p;
                        return j;
                    }
                    fail("expected ',' or ']'");
                }
            }
            if (c == '"')
            {
                j.type = String;
                j.s = parseString();
                return j;
            }
            if (c == '-' || (c >= '0' && c <= '9'))
                return parseNumber();
            if (end - p >= 4 && !strncmp(p,"true",4))
            {
                p += 4;
                j.type = Bool;
                j.b = true;
                return j;
            }
            if (end - p >= 5 && !strncmp(p,"false",5))
            {
                p += 5;
                j.type = Bool;
                j.b = false;
                return j;
            }
            if (end - p >= 4 && !strncmp(p,"null",4))
            {
                p += 4;
                return j;
            }
            fail(std::string("unexpected character '") + c + "'");
        }
    };

public:
    Json() {}

    static Json parse(const char *text, size_t len)
    {
        Parser ps{text,text+len};
        Json j = ps.parseValue(0);
        ps.skipWs();
        if (ps.p != ps.end)
            ps.fail("trailing characters");
        return j;
    }

    static Json parse(const std::string& text)
    {
        return parse(text.data(),text.size());
    }

    // predicates as in utils/json.cpp:28-66
    bool isNull() const { return type == Null; }
    bool isBool() const { return type == Bool; }
    bool isInt() const { return type == Int || type == UInt; }
    bool isFloat() const { return type == Int || type == UInt || type == Float; }
    bool isString() const { return type == String; }
    bool isArray() const { return type == Array; }
    bool isObject() const { return type == Object; }

    size_t size() const
    {
        if (isArray()) return arr.size();
        if (isObject()) return obj.size();
        throw JsonError("Json::size(): type is not an array or object");
    }

    bool boolValue() const
    {
        if (isBool()) return b;
        throw JsonError("Json::boolValue(): type is not bool");
    }

    int64_t intValue() const
    {
        if (type == Int) return i;
        if (type == UInt) return (int64_t)u;
        throw JsonError("Json::intValue(): type is not int");
    }

    double floatValue() const
    {
        if (type == Float) return f;
        if (type == Int) return (double)i;
        if (type == UInt) return (double)u;
        throw JsonError("Json::floatValue(): type is not float");
    }

    const std::string& stringValue() const
    {
        if (isString()) return s;
        throw JsonError("Json::stringValue(): type is not string");
    }

    const std::vector<Json>& arrayValue() const
    {
        if (isArray()) return arr;
        throw JsonError("Json::arrayValue(): type is not array");
    }

    const std::map<std::string,Json>& objectValue() const
    {
        if (isObject()) return obj;
        throw JsonError("Json::objectValue(): type is not object");
    }

    // valueAt (json.cpp:137-159): true and sets value if the key exists
    bool valueAt(const char *key, Json& value) const
    {
        if (!isObject())
            return false;
        auto it = obj.find(key);
        if (it == obj.end())
            return false;
        Json tmp = it->second; // value may alias *this
        value = tmp;
        return true;
    }

    bool has(const char *key) const
    {
        return isObject() && obj.find(key) != obj.end();
    }

    const Json& operator[](size_t index) const
    {
        if (!isArray())
            throw JsonError("Json::operator[](size_t): not an array");
        if (index >= arr.size())
            throw JsonError("Json::operator[](size_t): array index out of range: "
                + std::to_string(index));
        return arr[index];
    }

    const Json& operator[](const char *key) const
    {
        if (!isObject())
            throw JsonError("Json::operator[](char*): not an object");
        auto it = obj.find(key);
        if (it == obj.end())
            throw JsonError("Json::operator[](char*): key does not exist: "
                + std::string(key));
        return it->second;
    }

    // replace a member (used for the "size" override)
    void set(const char *key, const Json& v)
    {
        if (!isObject())
            throw JsonError("Json::set(): not an object");
        obj[key] = v;
    }

    // Compact serialisation as `os << nlohmann::json` prints it (the "flame:" echo of
    // ffr_buf.cpp:129, via utils/json.cpp:203-207): keys in std::map order, no whitespace,
    // integers in decimal, doubles as the shortest digits that round-trip laid out by
    // nlohmann's rule (plain notation for decimal exponents -4 < n <= 15 with a trailing
    // ".0" on whole numbers, d.ddde[+-]XX otherwise), non-finite as null.
    void dump(std::string& out) const
    {
        switch (type)
        {
        case Null: out += "null"; break;
        case Bool: out += b ? "true" : "false"; break;
        case Int: out += std::to_string(i); break;
        case UInt: out += std::to_string(u); break;
        case Float: dumpFloat(out,f); break;
        case String: dumpString(out,s); break;
        case Array:
            out += '[';
            for (size_t k = 0; k < arr.size(); ++k)
            {
                if (k) out += ',';
                arr[k].dump(out);
            }
            out += ']';
            break;
        case Object:
        {
            out += '{';
            bool first = true;
            for (auto& kv : obj)
            {
                if (!first) out += ',';
                first = false;
                dumpString(out,kv.first);
                out += ':';
                kv.second.dump(out);
            }
            out += '}';
            break;
        }
        }
    }

    static void dumpString(std::string& out, const std::string& v)
    {
        out += '"';
        for (unsigned char c : v)
        {
            switch (c)
            {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20)
                {
                    char esc[8];
                    snprintf(esc,sizeof(esc),"\\u%04x",(unsigned)c);
                    out += esc;
                }
                else
                    out += (char)c;
            }
        }
        out += '"';
    }

    static void dumpFloat(std::string& out, double v)
    {
        if (!std::isfinite(v))
        {
            out += "null";
            return;
        }
        if (v == 0.0)
        {
            out += std::signbit(v) ? "-0.0" : "0.0";
            return;
        }
        if (v < 0)
        {
            out += '-';
            v = -v;
        }
        // shortest round-trip digits d1 d2 .. dk and decimal exponent: v = 0.d1..dk * 10^n
        char sci[40];
        auto res = std::to_chars(sci,sci + sizeof(sci),v,std::chars_format::scientific);
        std::string t(sci,res.ptr);
        const size_t epos = t.find('e');
        std::string digits;
        for (size_t k = 0; k < epos; ++k)
            if (t[k] != '.')
                digits += t[k];
        const int n = atoi(t.c_str() + epos + 1) + 1;
        const int k = (int)digits.size();
        const int min_exp = -4, max_exp = 15;
        if (k <= n && n <= max_exp)
        {
            out += digits;
            out.append((size_t)(n - k),'0');
            out += ".0";
        }
        else if (0 < n && n <= max_exp)
        {
            out.append(digits,0,(size_t)n);
            out += '.';
            out.append(digits,(size_t)n,std::string::npos);
        }
        else if (min_exp < n && n <= 0)
        {
            out += "0.";
            out.append((size_t)(-n),'0');
            out += digits;
        }
        else
        {
            out += digits[0];
            if (k > 1)
            {
                out += '.';
                out.append(digits,1,std::string::npos);
            }
            out += 'e';
            int e = n - 1;
            out += e < 0 ? '-' : '+';
            if (e < 0) e = -e;
            char eb[16];
            snprintf(eb,sizeof(eb),"%02d",e);
            out += eb;
        }
    }

    static Json makeArrayOfInts(const uint64_t *v, size_t n)
    {
        Json j;
        j.type = Array;
        for (size_t k = 0; k < n; ++k)
        {
            Json e;
            e.type = Int;
            e.i = (int64_t)v[k];
            j.arr.push_back(e);
        }
        return j;
    }
};

} // namespace ffr
