// Levelled stdout logger with the interface of the reference's Logger
// (/root/reference/demonstrator/include/Logger.h:46-80): `Logger(INFO) << "text" << value;` prints one line,
// filtered by the global LOGCFG.level; LOGCFG.headers switches the coloured "[INFO ] " labels.
#ifndef MESHLESSHYDRO_LOGGER_H
#define MESHLESSHYDRO_LOGGER_H

#include <iostream>
#include <sstream>
#include <string>

enum typelog { DEBUG, INFO, WARN, ERROR };

struct structlog {
    bool headers = false;
    typelog level = WARN;
    int myRank = -1;     // rank label "(r)" in front of the header when >= 0 (multi-GPU launches)
    int outputRank = -1; // -1: every rank prints
};

extern structlog LOGCFG;

class Logger {
public:
    Logger() {}
    explicit Logger(typelog type);
    ~Logger();

    template <class T> Logger &operator<<(const T &msg) {
        if (enabled()) {
            line << msg;
            opened = true;
        }
        return *this;
    }

private:
    bool enabled() const {
        return msglevel >= LOGCFG.level && (msglevel >= ERROR || LOGCFG.outputRank == -1 || LOGCFG.myRank == LOGCFG.outputRank);
    }
    std::ostringstream line; // the whole line is emitted at once (ranks do not interleave inside a line)
    bool opened = false;
    typelog msglevel = DEBUG;
};

#endif // MESHLESSHYDRO_LOGGER_H
